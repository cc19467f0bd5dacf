"""BASELINE.json config 3: chunk-size sweep 4 KiB - 1 MiB over 1..N queue pairs, device-resident and
end-to-end from pinned host buffers (PCIe inside the timed region), 256 MiB of the lineitem-like workload.
Prints one JSON line per point.   usage: python tools/chunk_sweep.py [MiB] [qps,qps,...]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402


def run(data, seg, qps, pinned, reps=3, device_slots=False):
    L = capi.lib()
    U = data.size
    n = (U + seg - 1) // seg
    dev = CompressDevice(0, qps).Initialize(Configuration(
        decompressed_seg_size=seg, max_preallocate_memzones=n + 64,
        slot_mem_kind=capi.MEM_PINNED if pinned and not device_slots else capi.MEM_DEVICE))
    if pinned:
        h_in, h_out = C.c_void_p(), C.c_void_p()
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, U, 64, C.byref(h_in)))
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, n * seg, 64, C.byref(h_out)))
        C.memmove(h_in.value, data.ctypes.data, U)
        src_ptr, out_ptr = h_in.value, h_out.value
    else:
        src = torch.from_numpy(data).cuda()
        out = torch.empty(n * seg + 64, dtype=torch.uint8, device="cuda")
        src_ptr, out_ptr = src.data_ptr(), out.data_ptr()
    torch.cuda.synchronize()
    ops, slots = dev.compress_ops(src_ptr, U)
    per = (n + qps - 1) // qps
    parts = [(q * per, min(n, (q + 1) * per)) for q in range(qps) if q * per < n]
    best_c = best_d = 1e9
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        res = [dev.enqueue("deflate", q, ops[a:b]) for q, (a, b) in enumerate(parts)]
        for q in range(len(parts)):
            dev.wait(q)
        t1 = time.perf_counter()
        produced = np.concatenate([r["produced"] for r in res])
        pend = [dev.enqueue("inflate", q, dev.decompress_ops(slots[a:b], produced[a:b], out_ptr + a * seg))
                for q, (a, b) in enumerate(parts)]
        for q in range(len(parts)):
            dev.wait(q)
        t2 = time.perf_counter()
        best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    total = sum(int(r["produced"].sum()) for r in pend)
    if pinned:
        back = np.ctypeslib.as_array(C.cast(out_ptr, C.POINTER(C.c_uint8)), shape=(U,))
        ok = total == U and np.array_equal(back, data)
        for b in (h_in, h_out):
            L.bitar_mem_free(capi.MEM_PINNED, 0, b)
    else:
        ok = total == U and bool(torch.equal(out[:U], src))
    for s in slots[::-1]:
        dev.put_slot(s)
    dev.close()
    return {"seg": seg, "qps": qps, "buffers": ("pinned_host+device_slots" if device_slots else "pinned_host") if pinned else "device", "chunks": n,
            "compress_gbps": round(U / best_c / 1e9, 2), "decompress_gbps": round(U / best_d / 1e9, 2),
            "ratio": round(U / float(produced.sum()), 3), "ok": ok}


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    qps_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]
    data = synth.lineitem_like(mib << 20)
    # SWEEP_BUFFERS=device|pinned and SWEEP_SEGS=4096,65536,... restrict the sweep
    only = os.environ.get("SWEEP_BUFFERS", "")
    segs = [int(x) for x in os.environ.get("SWEEP_SEGS", "").split(",") if x] or \
        [4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576]
    if only == "mixed":    # experiment: input / output buffers pinned, compressed slots in device memory
        for seg in segs:
            for qps in qps_list:
                print(json.dumps(run(data, seg, qps, True, device_slots=True)), flush=True)
        return
    for pinned in (False, True):
        if only and only != ("pinned" if pinned else "device"):
            continue
        for seg in segs:
            for qps in qps_list:
                print(json.dumps(run(data, seg, qps, pinned)), flush=True)


if __name__ == "__main__":
    main()
