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
import gc
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
REF_SAMPLE_MIB = 96          # --impl reference: MiB of the workload one step processes (equal parts of the three columns)


def dram_traffic():
    """DRAM bytes per uncompressed byte of the two codec kernels, as tools/profile_round.sh wrote them from an
    `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def config_of(args):
    """The workload both arms run (identical in both lines, so that the driver's same_config holds)."""
    n = ((args.mib << 20) + args.seg - 1) // args.seg
    return {"workload": f"{args.mib} MiB/GPU lineitem-like Arrow columns (sorted int64, dict int32, f64 prices), seg {args.seg} B "
                        f"({n} chunks), dynamic Huffman, level-1-equivalent, Compress() then Decompress()",
            "seg": args.seg, "chunks_per_gpu": n, "mib_per_gpu": args.mib,
            "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no flush needed",
            "reference_sample": f"a step of --impl reference runs the first {REF_SAMPLE_MIB // 3} MiB of each column third of this buffer "
                                f"({REF_SAMPLE_MIB} MiB) on all host threads; its throughput is per byte of that sample"}


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


def cpu_baseline(data, seg, threads):
    """Times the CPU path (oracle: bitar chunking over zlib level 1, one z_stream per worker) on a bounded sample of
    the workload (the same sample --impl reference uses).  Returns the cpu_baseline record."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    sample = column_sample(data, seg, REF_SAMPLE_MIB)
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
    }


def column_sample(data, seg, mib):
    """Equal parts of the three column thirds, `mib` MiB in all (whole segments)."""
    third = data.size // 3
    part = min(third, (mib << 20) // 3) // seg * seg
    return np.concatenate([data[i * third:i * third + part] for i in range(3)])


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (bitar chunking + zlib, restated in oracle/ because
    the reference cannot be built here) on all host threads; every step is the same bounded sample of the workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from bitar_b200 import synth
    threads = max(1, (os.cpu_count() or 2) - 1)     # one core stays the main lcore, src/driver.cc:198-220
    data = synth.lineitem_like(args.mib << 20)
    sample = column_sample(data, args.seg, REF_SAMPLE_MIB)
    del data
    times, td, ti, produced = [], [], [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        slots, produced = O.compress_buffer(sample, args.seg, threads=threads)
        t1 = time.perf_counter()
        out, got = O.decompress_buffer(slots, produced, args.seg, threads=threads)
        t2 = time.perf_counter()
        if i == 0:
            assert out.size == sample.size and np.array_equal(out, sample)
        if i >= args.warmup:
            times.append(t2 - t0)
            td.append(t1 - t0)
            ti.append(t2 - t1)
    u = sample.size
    value = u / float(np.mean(times)) / 1e9          # mean over the timed steps, the timer ms_per_step comes from
    base = {"value": value, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"{u >> 20} MiB of the workload (equal column mix), seg {args.seg}, zlib {O.lib().oracle_zlib_version().decode()} "
                      f"raw deflate level 1 / inflate, one reused z_stream per worker",
            "deflate_gbps": u / float(np.mean(td)) / 1e9, "inflate_gbps": u / float(np.mean(ti)) / 1e9,
            "ratio": u / float(produced.sum())}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(args), "timing": "host wall clock around Compress() + Decompress() of the sample, mean over the timed steps",
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
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE config 4 leg (8 GiB per GPU, generated on the device)")
    ap.add_argument("--config4-gib", type=int, default=8)
    ap.add_argument("--no-extras", action="store_true", help="skip the ratio corpus and the foreign-stream inflate leg")
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
        Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 256))
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

    # ---------------- ratio corpus and reference-compressed streams (rank 0) ----------------
    ratio_corpus = foreign = None
    if not args.no_extras and rank == 0:
        ratio_corpus = run_ratio_corpus(dev, capi, torch, seg)
        foreign = run_foreign(dev, torch, data, seg, mib=args.mib)   # the whole workload: fewer streams than lanes of the kernel grid would understate it

    # ---------------- BASELINE config 4: 8 GiB per GPU, generated on the device ----------------
    config4 = None
    if not args.no_config4:
        config4 = run_config4(args, capi, torch, dist, world, local_rank, rank, seg)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    base = None
    if not args.no_cpu and world == 1 and rank == 0:
        threads = max(1, (os.cpu_count() or 2) - 1)
        base = cpu_baseline(data, seg, threads)

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
        traffic = dram_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(args), "timing": "device-resident: CUDA events on the queue pair's stream, max over ranks",
            "deflate_gbps": world * U / (td_ms * 1e-3) / 1e9, "inflate_gbps": world * U / (ti_ms * 1e-3) / 1e9,
            "deflate_kernel_ms": kd_ms, "inflate_kernel_ms": ki_ms, "wall_ms_per_step": wall_ms,
            "ratio": U / Cbytes, "zlib_level1_ratio": zratio,
            # dominant kernel = deflate_kernel (the larger share of the device-resident step's GPU time, see the ncu launch
            # list under profiles/).  achieved = algorithmic bytes (U + C) / the kernel's time from CUDA events in this run;
            # traffic = DRAM bytes of one `ncu --set full` capture, scaled to this launch's bytes (profiles/r02_dram_traffic.json).
            "roofline": {"bound": "hbm", "kernel": "deflate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": traffic["deflate_bytes_per_byte"] * U if traffic else None, "peak_source": which,
                         "traffic_source": traffic["source"] if traffic else None,
                         "algorithmic_bytes": "U + C per launch (read input once, write the stream once)",
                         "note": "issue-bound integer kernel (ncu: issue-slot utilisation, not DRAM, is what saturates)",
                         "inflate": {"kernel": "inflate_tok_kernel", "achieved": (U + Cbytes) / (ki_ms * 1e-3) / 1e9,
                                     "frac": (U + Cbytes) / (ki_ms * 1e-3) / 1e9 / peak,
                                     "traffic": traffic["inflate_bytes_per_byte"] * U if traffic else None}},
            "clocks": clocks, "gpu_launches": launches,
        }
        if ratio_corpus:
            line["ratio_corpus"] = ratio_corpus
        if foreign:
            line.update(foreign)
        if config4:
            line["config4"] = config4
        if e2e:
            line["e2e"] = e2e
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ratio_corpus(dev, capi, torch, seg, nbytes=4 << 20):
    """Compressed size next to zlib level 1 on every input of synth.ratio_corpus() (identical chunks)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from bitar_b200 import synth
    out = {}
    threads = max(1, (os.cpu_count() or 2) - 1)
    for name, d in synth.ratio_corpus(nbytes).items():
        src = torch.from_numpy(d).cuda()
        torch.cuda.synchronize()
        ops, slots = dev.compress_ops(src.data_ptr(), d.size)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        dev.put_slots(slots)
        _, zp = O.compress_buffer(d, seg, threads=threads)
        g, z = int(res["produced"].sum()), int(zp.sum())
        out[name] = {"ratio": round(d.size / g, 4), "zlib_level1_ratio": round(d.size / z, 4), "bytes_vs_zlib": round(g / z, 4)}
    return out


def run_foreign(dev, torch, data, seg, mib=256):
    """Inflate of reference-compressed streams (zlib level 1, no index: the reference's own output,
    /root/reference/src/memory.cc:432-505) of the same workload, device-resident."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    sample = column_sample(data, seg, mib)
    slots, produced = O.compress_buffer(sample, seg, threads=max(1, (os.cpu_count() or 2) - 1))
    d_slots = torch.from_numpy(slots.reshape(-1)).cuda()
    out = torch.empty(slots.shape[0] * seg + 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ptrs = np.uint64(d_slots.data_ptr()) + np.arange(slots.shape[0], dtype=np.uint64) * np.uint64(slots.shape[1])
    iops = dev.decompress_ops(ptrs, produced, out.data_ptr())
    best = 1e9
    for _ in range(3):
        r = dev.enqueue("inflate", 0, iops)
        dev.wait(0)
        best = min(best, dev.last_ms(0)[1])
    ok = int(r["produced"].sum()) == sample.size and bool(torch.equal(out[:sample.size], torch.from_numpy(sample).cuda()))
    peak, _ = peaks()
    algo = (sample.size + float(produced.sum())) / (best * 1e-3) / 1e9       # read the stream once, write the output once
    return {"inflate_foreign_gbps": sample.size / (best * 1e-3) / 1e9, "inflate_foreign_ok": ok,
            "inflate_foreign_roofline": {"kernel": "inflate_spec_kernel", "bound": "hbm", "achieved": algo, "peak": peak, "unit": "GB/s",
                                         "frac": algo / peak, "zlib_level1_ratio": sample.size / float(produced.sum())},
            "inflate_foreign_sample": f"{sample.size >> 20} MiB of the workload compressed by zlib level 1 (no index): speculative lane-parallel kernel (inflate_spec_kernel.cuh)"}


def gen_lineitem_on_device(torch, nbytes, seed):
    """synth.lineitem_like() generated on the device (BASELINE config 4: the 64 GiB corpus does not fit the host)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    third = (nbytes // 3) // 8 * 8
    a = torch.cumsum(torch.randint(0, 4, (third // 8,), generator=g, device="cuda", dtype=torch.int64), 0)
    p = torch.tensor([.30, .20, .15, .12, .10, .08, .05], device="cuda").cumsum(0)
    b = torch.bucketize(torch.rand(third // 4, generator=g, device="cuda"), p).clamp_(max=6).to(torch.int32)
    rest = (nbytes - 2 * third) // 8
    c = torch.randint(90_000, 10_500_000, (rest,), generator=g, device="cuda", dtype=torch.int64).to(torch.float64) / 100.0
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    buf[:third] = a.view(torch.uint8)
    buf[third:2 * third] = b.view(torch.uint8)
    buf[2 * third:2 * third + rest * 8] = c.view(torch.uint8)
    return buf


def run_config4(args, capi, torch, dist, world, local_rank, rank, seg):
    """BASELINE config 4: a 64 GiB columnar corpus sharded 8 GiB per GPU, compress + inflate, device-resident: eight
    1 GiB buffers per GPU, one per queue pair, all queue pairs at once; the pool holds ceil(bytes / S) + n_qp slots as
    the reference sizes it (/root/reference/apps/app_common.cc:94-100)."""
    from bitar_b200.engine import CompressDevice, Configuration
    gib = args.config4_gib
    per = 1 << 30
    n_per = (per + seg - 1) // seg
    dev = CompressDevice(local_rank, gib).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=gib * n_per + gib))
    bufs = [gen_lineitem_on_device(torch, per, 20261018 + 1000 * rank + k) for k in range(gib)]
    outs = [torch.empty(n_per * seg + 64, dtype=torch.uint8, device="cuda") for _ in range(gib)]
    torch.cuda.synchronize()
    slots_free0 = dev.slots_free()

    def one_pass():
        t0 = time.perf_counter()
        plan = [dev.compress_ops(b.data_ptr(), per) for b in bufs]          # Take() of every chunk's slot
        res = [dev.enqueue("deflate", q, plan[q][0]) for q in range(gib)]
        for q in range(gib):
            dev.wait(q)
        t1 = time.perf_counter()
        pend = [dev.enqueue("inflate", q, dev.decompress_ops(plan[q][1], res[q]["produced"], outs[q].data_ptr())) for q in range(gib)]
        for q in range(gib):
            dev.wait(q)
        t2 = time.perf_counter()
        comp = sum(int(r["produced"].sum()) for r in res)
        total = sum(int(r["produced"].sum()) for r in pend)
        for q in range(gib):
            dev.put_slots(plan[q][1])
        return t1 - t0, t2 - t1, comp, total

    one_pass()                                                               # warm-up (scratch allocation, clocks)
    td, ti, comp, total = one_pass()
    ok = total == gib * per and all(bool(torch.equal(o[:per], b)) for o, b in zip(outs, bufs))
    t = torch.tensor([td, ti], dtype=torch.float64, device="cuda")
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    td, ti = [float(x) for x in t.cpu()]
    rec = {"gib_per_gpu": gib, "chunks_per_gpu": gib * n_per, "pool_slots": gib * n_per + gib, "queue_pairs": gib,
           "deflate_gbps": world * gib * per / td / 1e9, "inflate_gbps": world * gib * per / ti / 1e9,
           "round_trip_gbps": world * gib * per / (td + ti) / 1e9, "ratio": gib * per / comp, "round_trip_ok": bool(int(flag.cpu()[0])),
           "slots_back": dev.slots_free() == slots_free0,
           "timing": "host wall clock around all queue pairs of a device (Take + enqueue + wait), max over ranks; data generated on the device"}
    dev.close()
    del bufs, outs
    torch.cuda.empty_cache()
    return rec


def run_e2e(args, L, capi, C, torch, dist, world, local_rank, data, seg, n, U):
    """Compress() + Decompress() with every user-visible buffer in pinned host memory, as a bitar
    application holds them (the Rtememzone pool of apps/demo_app.cc:119-122,517-522):
      compress  : the kernel pulls each chunk from host memory with its TMA bulk copy (H2D = U bytes) and
                  writes the streams straight into pinned output slots (D2H = C bytes) -- zero-copy;
      decompress: called with the pinned slots and the pinned destination; the library gathers the slots into
                  device memory (H2D = C bytes), inflates, and scatters the result back (D2H = U bytes).
    The input is split evenly over --qps queue pairs that run concurrently (apps/demo_app.cc:577-596).  Every step
    Take()s its output slots, builds its op lists and Recycle()s the slots inside the timed region, as the reference's
    Compress() / Recycle() do per chunk (src/memory.cc:405-425, src/device.cc:320-327).
    Two schedules are timed: `pipelined` (the headline e2e value) chains CompressAsync -> callback -> DecompressAsync per
    queue pair (src/include/util.h:216-236), the odd queue pairs half a phase behind the even ones, so that one half
    decompresses (device-to-host traffic) while the other compresses (host-to-device traffic); `phase_separated` waits
    for every queue pair's compress before any decompress starts, as demo_app's evaluation does."""
    from collections import deque

    from bitar_b200.engine import CompressDevice, Configuration
    qps = max(1, args.qps)
    dev = CompressDevice(local_rank, qps).Initialize(
        Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 64, slot_mem_kind=capi.MEM_PINNED))
    h_in, h_out = C.c_void_p(), C.c_void_p()
    capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, local_rank, U, 64, C.byref(h_in)))
    capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, local_rank, n * seg, 64, C.byref(h_out)))
    C.memmove(h_in.value, data.ctypes.data, U)
    torch.cuda.synchronize()
    per = (n + qps - 1) // qps
    parts = [(q * per, min(n, (q + 1) * per)) for q in range(qps) if q * per < n]

    def part_ops(a, b):
        nbytes = min(U, b * seg) - a * seg
        return dev.compress_ops(h_in.value + a * seg, nbytes)             # Take()s b - a slots

    phase_ms = [0.0, 0.0]

    def step_separated():
        t_a = time.perf_counter()
        plan = [part_ops(a, b) for a, b in parts]
        results = [dev.enqueue("deflate", q, plan[q][0]) for q in range(len(parts))]
        for q in range(len(parts)):
            dev.wait(q)
        phase_ms[0] = (time.perf_counter() - t_a) * 1e3
        pending = []
        for q, (a, b) in enumerate(parts):
            iops = dev.decompress_ops(plan[q][1], results[q]["produced"], h_out.value + a * seg)
            pending.append(dev.enqueue("inflate", q, iops))
        for q in range(len(parts)):
            dev.wait(q)
        comp = sum(int(r["produced"].sum()) for r in results)
        total = sum(int(r["produced"].sum()) for r in pending)
        for q in range(len(parts)):
            dev.put_slots(plan[q][1])
        phase_ms[1] = (time.perf_counter() - t_a) * 1e3 - phase_ms[0]
        return comp, total

    host_share = [0.0, 0.0]                                                 # the driving thread: seconds asleep (nothing to do) / seconds in the loop
    K = 4                                                                   # sub-parts per queue pair (2: 42 ms, 4: 39 ms, 8: 44 ms per GiB)

    def step_pipelined():
        todo = []
        for a, b in parts:
            m = (b - a + K - 1) // K
            todo.append(deque((a + k * m, min(b, a + (k + 1) * m)) for k in range(K) if a + k * m < b))
        state = ["idle"] * len(parts)                                        # idle -> c -> d -> idle
        cur = [None] * len(parts)
        comp = total = 0
        odd_released = len(parts) < 2

        def start_compress(q):
            a, b = todo[q].popleft()
            ops_, slots_ = part_ops(a, b)
            cur[q] = (a, b, slots_, dev.enqueue("deflate", q, ops_))
            state[q] = "c"

        for q in range(0, len(parts), 2):
            start_compress(q)
        done = 0
        t_busy0 = time.perf_counter()
        idle_s = 0.0
        while done < len(parts):
            progressed = False
            for q in range(len(parts)):
                if state[q] in ("c", "d") and not dev.busy(q):
                    capi.check(L.bitar_qp_result(dev._h, q))
                    a, b, slots_, res_ = cur[q]
                    if state[q] == "c":
                        comp += int(res_["produced"].sum())
                        iops = dev.decompress_ops(slots_, res_["produced"], h_out.value + a * seg)
                        cur[q] = (a, b, slots_, dev.enqueue("inflate", q, iops))
                        state[q] = "d"
                        if not odd_released:                                 # the odd queue pairs start half a phase late
                            odd_released = True
                            for o in range(1, len(parts), 2):
                                start_compress(o)
                    else:
                        total += int(res_["produced"].sum())
                        dev.put_slots(slots_)
                        if todo[q]:
                            start_compress(q)
                        else:
                            state[q] = "idle"
                            done += 1
                    progressed = True
            if not progressed:
                t_i = time.perf_counter()
                time.sleep(0.00005)
                idle_s += time.perf_counter() - t_i
        host_share[0] += idle_s
        host_share[1] += time.perf_counter() - t_busy0
        return comp, total

    e2e_steps = max(args.steps, 12)           # the mean over at least 12 steps: a step is ~40 ms of host-driven scheduling, and one hiccup of the box (a 70 ms step among six was seen) should not decide the number

    def timed(step):
        for _ in range(max(1, args.warmup)):
            comp, total = step()
        back = np.ctypeslib.as_array(C.cast(h_out.value, C.POINTER(C.c_uint8)), shape=(U,))
        assert total == U and np.array_equal(back, data), "end-to-end round trip differs"
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()                          # (a collection inside a 40 ms step showed up as a 65 ms step)
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            t_step = time.perf_counter()
            comp, total = step()
            step_ms.append(round((time.perf_counter() - t_step) * 1e3, 3))
        torch.cuda.synchronize()
        gc.enable()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.cpu()[0]), step_ms, comp

    dt_sep, ms_sep, cbytes = timed(step_separated)
    host_share[0] = host_share[1] = 0.0
    sep_phases = {"compress_ms": round(phase_ms[0], 3), "decompress_ms": round(phase_ms[1], 3)}
    dt, step_ms, cbytes = timed(step_pipelined)
    # PCIe roofline of this leg, measured in the same run: plain pinned <-> device copies of U bytes, one direction at
    # a time and both at once (two queue pairs' streams)
    d_tmp = torch.empty(U, dtype=torch.uint8, device="cuda")
    d_tmp2 = torch.empty(U, dtype=torch.uint8, device="cuda")
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
    if len(parts) > 1:
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            capi.check(L.bitar_qp_memcpy(dev._h, 0, d_tmp2.data_ptr(), h_in.value, U))
            capi.check(L.bitar_qp_memcpy(dev._h, 1, h_out.value, d_tmp.data_ptr(), U))
            dev.wait(0)
            dev.wait(1)
            best = min(best, time.perf_counter() - t1)
        pcie["both_directions_gbps_each"] = U / best / 1e9
    del d_tmp, d_tmp2
    # each step moves U + C bytes in each direction; with both directions busy at once the floor is (U + C) / bandwidth
    both = pcie.get("both_directions_gbps_each", min(pcie["h2d_gbps"], pcie["d2h_gbps"]))
    floor_pipe = (U + cbytes) / (both * 1e9)
    floor_sep = U / (pcie["h2d_gbps"] * 1e9) + U / (pcie["d2h_gbps"] * 1e9)
    pcie["frac_of_pcie_floor"] = floor_pipe / dt
    pcie["phase_separated_frac_of_pcie_floor"] = floor_sep / dt_sep
    for b in (h_in, h_out):
        capi.check(L.bitar_mem_free(capi.MEM_PINNED, local_rank, b))
    dev.close()
    return {"value": world * U / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(U + cbytes),
            "d2h_bytes_per_step": int(cbytes + U), "ms_per_step": dt * 1e3, "median_ms_per_step": float(np.median(step_ms)), "steps": e2e_steps, "step_ms": step_ms, "queue_pairs": len(parts),
            "schedule": f"pipelined: per queue pair Compress -> Decompress of {K} sub-parts back to back, odd queue pairs half a phase behind",
            "host_thread_idle_frac": round(host_share[0] / host_share[1], 3) if host_share[1] else None,
            "phase_separated": {"value": world * U / dt_sep / 1e9, "ms_per_step": dt_sep * 1e3, "step_ms": ms_sep, "last_step": sep_phases}, "pcie": pcie,
            "path": "pinned host in/out through the C-ABI: compress reads and writes host memory in place (zero-copy over "
                    "PCIe), decompress is staged through device memory by the library in batches on three extra streams per queue pair (strided copy-engine gather -- rows go at the pitch of the batch's widest stream, so somewhat more than C bytes cross PCIe --, inflate, copy-engine copy-back); Take / op lists / Recycle inside the timed step"}


if __name__ == "__main__":
    main()
