#!/usr/bin/env python
"""bench.py -- the reference's headline measurement on B200: chunked raw-DEFLATE compress + inflate of
an Arrow-columnar buffer behind bitar's API (BASELINE.json), device-resident and end-to-end.

A "step" is one pass of the hot path over one batch, exactly what bitar's demo_app times
(/root/reference/apps/demo_app.cc:487-548): Compress() the whole buffer, then Decompress() it back.
    value      = uncompressed bytes / (t_compress + t_decompress), inputs resident in HBM   [GB/s]
    e2e.value  = the same through the C-ABI with HOST (pinned) buffers, PCIe inside the timed region
Sub-metrics (deflate / inflate GB/s, compression ratio next to zlib level 1) ride along in the JSON.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mib M] [--seg S]
N > 1 is launched by torchrun (one rank per GPU); chunks shard across ranks with no data-path
collective (weak scaling: every rank owns its own buffer), NCCL only gathers the timing.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deflate+inflate round-trip throughput, uncompressed bytes (bitar Compress->Decompress)"
SEG_DEFAULT = 59460          # apps/app_common.h:39 kDecompressedSegSize
# DRAM bytes moved per uncompressed byte, from ncu (dram__bytes_read.sum + dram__bytes_write.sum, 256 MiB launch)
DEFLATE_DRAM_BYTES_PER_BYTE = (289.780736e6 + 109.076992e6) / 268435456
INFLATE_DRAM_BYTES_PER_BYTE = (161.789440e6 + 255.810048e6) / 268435456


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(data, seg, threads, want_seconds=10.0):
    """Times the CPU path (oracle: bitar chunking over zlib level 1, one z_stream per worker) on a
    bounded sample of the workload.  Returns (dict, sample bytes)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    sample = data
    # ~0.1 GB/s/core deflate: bound the sample so one pass takes roughly want_seconds / 3
    budget = int(0.08e9 * threads * want_seconds / 3)
    budget = max(seg * threads * 4, budget // seg * seg)
    if sample.size > budget:
        # equal parts of each third so the column mix is preserved
        third = data.size // 3
        part = budget // 3 // seg * seg
        sample = np.concatenate([data[i * third:i * third + part] for i in range(3)])
    t0 = time.perf_counter()
    slots, produced = O.compress_buffer(sample, seg, threads=threads)
    t1 = time.perf_counter()
    out, got = O.decompress_buffer(slots, produced, seg, threads=threads)
    t2 = time.perf_counter()
    assert out.size == sample.size and np.array_equal(out, sample)
    u = sample.size
    return {
        "value": u / (t2 - t0) / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
        "sample": f"{u >> 20} MiB of the same workload (equal column mix), seg {seg}, zlib {O.lib().oracle_zlib_version().decode()} "
                  f"raw deflate level 1 / inflate, one reused z_stream per worker",
        "deflate_gbps": u / (t1 - t0) / 1e9, "inflate_gbps": u / (t2 - t1) / 1e9,
        "ratio": u / float(produced.sum()),
    }, u


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (bitar chunking + zlib, restated in oracle/ because
    the reference cannot be built here) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from bitar_b200 import synth
    threads = max(1, (os.cpu_count() or 2) - 1)     # one core stays the main lcore, src/driver.cc:198-220
    data = synth.lineitem_like(min(args.mib, 256) << 20)
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        base, u = cpu_baseline(data, args.seg, threads, want_seconds=4.0)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    value = base["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.mib} MiB/GPU lineitem-like Arrow columns (sorted int64, dict int32, f64 prices), "
                               f"seg {args.seg} B ({((args.mib << 20) + args.seg - 1) // args.seg} chunks), dynamic Huffman, level-1-equivalent",
                   "seg": args.seg, "chunks_per_gpu": ((args.mib << 20) + args.seg - 1) // args.seg,
                   "timing": f"host wall clock; every step is a bounded sample of that workload ({base['sample']})"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mib", type=int, default=1024, help="uncompressed MiB per GPU (BASELINE config 2: 1 GiB)")
    ap.add_argument("--seg", type=int, default=SEG_DEFAULT)
    ap.add_argument("--qps", type=int, default=8, help="queue pairs used by the end-to-end leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    from bitar_b200 import _capi as capi
    from bitar_b200 import synth
    from bitar_b200.engine import CompressDevice, Configuration

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = capi.lib()

    seg, U = args.seg, args.mib << 20
    data = synth.lineitem_like(U, seed=synth.SEED + rank)      # every rank owns its own shard
    n = (U + seg - 1) // seg

    # ---------------- device-resident leg ----------------
    dev = CompressDevice(local_rank, max(1, args.qps)).Initialize(
        Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 64))
    src = torch.from_numpy(data).cuda()
    out = torch.empty(n * seg + 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ops, slots = dev.compress_ops(src.data_ptr(), U)

    def step():
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        kd, td = dev.last_ms(0)
        iops = dev.decompress_ops(slots, res["produced"], out.data_ptr())
        ires = dev.enqueue("inflate", 0, iops)
        dev.wait(0)
        ki, ti = dev.last_ms(0)
        return res, ires, kd, td, ki, ti

    for _ in range(args.warmup):
        res, ires, *_ = step()
    assert int(ires["produced"].sum()) == U and bool(torch.equal(out[:U], src)), "round trip differs"
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kd_sum = td_sum = ki_sum = ti_sum = 0.0
    t_wall0 = time.perf_counter()
    launches0 = L.bitar_kernel_launches()
    for _ in range(args.steps):
        res, ires, kd, td, ki, ti = step()
        kd_sum, td_sum, ki_sum, ti_sum = kd_sum + kd, td_sum + td, ki_sum + ki, ti_sum + ti
    torch.cuda.synchronize()
    launches = int(L.bitar_kernel_launches() - launches0)   # this library's kernels inside the timed region
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    Cbytes = int(res["produced"].sum())
    dev_ms = (td_sum + ti_sum) / args.steps              # CUDA events on the queue pair's stream
    times = torch.tensor([dev_ms, td_sum / args.steps, ti_sum / args.steps, kd_sum / args.steps, ki_sum / args.steps,
                          1e3 * t_wall / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)     # max over ranks
    dev_ms, td_ms, ti_ms, kd_ms, ki_ms, wall_ms = [float(x) for x in times.cpu()]
    value = world * U / (dev_ms * 1e-3) / 1e9

    # ---------------- end-to-end leg: host (pinned) buffers through the C-ABI ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, L, capi, C, torch, dist, world, local_rank, data, seg, n, U)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    base = None
    if not args.no_cpu and world == 1 and rank == 0:
        threads = max(1, (os.cpu_count() or 2) - 1)
        base, _ = cpu_baseline(data, seg, threads)

    # zlib level-1 ratio on identical chunks (sample)
    zratio = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        third, part = U // 3, min(U // 3, 16 << 20) // seg * seg
        zs = np.concatenate([data[i * third:i * third + part] for i in range(3)])
        _, zp = O.compress_buffer(zs, seg, threads=max(1, (os.cpu_count() or 2) - 1))
        zratio = zs.size / float(zp.sum())

    dev.close()
    if rank == 0:
        peak, which = peaks()
        achieved = (U + Cbytes) / (kd_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.mib} MiB/GPU lineitem-like Arrow columns (sorted int64, dict int32, f64 prices), "
                                   f"seg {seg} B ({n} chunks), dynamic Huffman, level-1-equivalent, device-resident",
                       "seg": seg, "chunks_per_gpu": n, "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the queue pair's stream, max over ranks"},
            "deflate_gbps": world * U / (td_ms * 1e-3) / 1e9, "inflate_gbps": world * U / (ti_ms * 1e-3) / 1e9,
            "deflate_kernel_ms": kd_ms, "inflate_kernel_ms": ki_ms, "wall_ms_per_step": wall_ms,
            "ratio": U / Cbytes, "zlib_level1_ratio": zratio,
            # dominant kernel = deflate_kernel (74 % of the device-resident step's GPU time, profiles/r01_launches_bench_1GiB_q.csv).
            # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at 256 MiB
            # (profiles/r01_ncu_full_deflate_and_indexed_inflate_256MiB_q.txt), scaled linearly to this launch's bytes.
            "roofline": {"bound": "hbm", "kernel": "deflate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": DEFLATE_DRAM_BYTES_PER_BYTE * U, "peak_source": which,
                         "algorithmic_bytes": "U + C per launch (read input once, write the stream once)",
                         "note": "issue- and latency-bound integer kernel: 66 % issue-slot utilisation, 1.5 % DRAM throughput (ncu)",
                         "inflate": {"kernel": "inflate_indexed_kernel", "achieved": (U + Cbytes) / (ki_ms * 1e-3) / 1e9,
                                     "frac": (U + Cbytes) / (ki_ms * 1e-3) / 1e9 / peak,
                                     "traffic": INFLATE_DRAM_BYTES_PER_BYTE * U}},
            "clocks": clocks, "gpu_launches": launches,
        }
        if e2e:
            line["e2e"] = e2e
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, L, capi, C, torch, dist, world, local_rank, data, seg, n, U):
    """Compress() + Decompress() with every user-visible buffer in pinned host memory, as a bitar
    application holds them (the Rtememzone pool of apps/demo_app.cc:119-122,517-522):
      compress  : the kernel pulls each chunk from host memory with its TMA bulk copy (H2D = U bytes) and
                  writes the streams straight into pinned output slots (D2H = C bytes) -- zero-copy;
      decompress: called with the pinned slots and the pinned destination; the library gathers the slots into
                  device memory (H2D = C bytes), inflates, and scatters the result back (D2H = U bytes).
    The input is split evenly over --qps queue pairs that run concurrently (apps/demo_app.cc:577-596)."""
    from bitar_b200.engine import CompressDevice, Configuration
    qps = max(1, args.qps)
    dev = CompressDevice(local_rank, qps).Initialize(
        Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 64, slot_mem_kind=capi.MEM_PINNED))
    h_in, h_out = C.c_void_p(), C.c_void_p()
    capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, local_rank, U, 64, C.byref(h_in)))
    capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, local_rank, n * seg, 64, C.byref(h_out)))
    C.memmove(h_in.value, data.ctypes.data, U)
    torch.cuda.synchronize()
    ops, slots = dev.compress_ops(h_in.value, U)
    per = (n + qps - 1) // qps
    parts = [(q * per, min(n, (q + 1) * per)) for q in range(qps) if q * per < n]

    def step():
        # Compress(): every queue pair deflates its part, reading the pinned input and writing the pinned slots
        results = [dev.enqueue("deflate", q, ops[a:b]) for q, (a, b) in enumerate(parts)]
        for q in range(len(parts)):
            dev.wait(q)
        produced = np.concatenate([r["produced"] for r in results])
        # Decompress(): pinned slots -> pinned destination; the library stages both through device memory
        pending = []
        for q, (a, b) in enumerate(parts):
            iops = dev.decompress_ops(slots[a:b], produced[a:b], h_out.value + a * seg)
            pending.append(dev.enqueue("inflate", q, iops))
        for q in range(len(parts)):
            dev.wait(q)
        total = sum(int(r["produced"].sum()) for r in pending)
        return int(produced.sum()), int(produced.sum()), total

    for _ in range(max(1, args.warmup)):
        cbytes, h2d, total = step()
    back = np.ctypeslib.as_array(C.cast(h_out.value, C.POINTER(C.c_uint8)), shape=(U,))
    assert total == U and np.array_equal(back, data), "end-to-end round trip differs"
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        t_step = time.perf_counter()
        cbytes, h2d, total = step()
        step_ms.append(round((time.perf_counter() - t_step) * 1e3, 3))
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.cpu()[0])
    # PCIe roofline of this leg, measured in the same run: plain pinned <-> device copies of U bytes
    d_tmp = torch.empty(U, dtype=torch.uint8, device="cuda")
    pcie = {}
    for name, dst, src in (("h2d_gbps", d_tmp.data_ptr(), h_in.value), ("d2h_gbps", h_out.value, d_tmp.data_ptr())):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            capi.check(L.bitar_qp_memcpy(dev._h, 0, dst, src, U))
            dev.wait(0)
            best = min(best, time.perf_counter() - t1)
        pcie[name] = U / best / 1e9
    del d_tmp
    # each step moves U + C bytes in each direction; with both directions overlapped the floor is max(dir) / bw
    floor = max((U + h2d) / (pcie["h2d_gbps"] * 1e9), (cbytes + U) / (pcie["d2h_gbps"] * 1e9))
    pcie["frac_of_pcie_floor"] = floor / dt
    for b in (h_in, h_out):
        capi.check(L.bitar_mem_free(capi.MEM_PINNED, local_rank, b))
    for s in slots[::-1]:
        dev.put_slot(s)
    dev.close()
    return {"value": world * U / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(U + h2d),
            "d2h_bytes_per_step": int(cbytes + U), "ms_per_step": dt * 1e3, "step_ms": step_ms, "queue_pairs": len(parts), "pcie": pcie,
            "path": "pinned host in/out through the C-ABI: compress reads and writes host memory in place (zero-copy over "
                    "PCIe), decompress is staged through device memory by the library in batches on three extra streams per queue pair (strided copy-engine gather -- rows go at the pitch of the batch's widest stream, so somewhat more than C bytes cross PCIe --, inflate, copy-engine copy-back)"}


if __name__ == "__main__":
    main()
