"""Probe: does a library call enqueued AFTER a large host-to-device copy on another stream wait for it?  (It did while
the descriptor upload went through the copy engine, which serves requests in order across all streams.)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitar_b200 import _capi as capi, synth
from bitar_b200.engine import CompressDevice, Configuration
seg = 59460
data = synth.lineitem_like(1 << 30)
n = (data.size + seg - 1) // seg
dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 8))
src = torch.from_numpy(data).cuda()
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
ops, slots = dev.compress_ops(src.data_ptr(), data.size)
def run(order):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for o in order:
        if o == "k": dev.enqueue("deflate", 0, ops)
        if o == "c":
            with torch.cuda.stream(s2): d.copy_(h, non_blocking=True)
        if o == "c8":
            with torch.cuda.stream(s2):
                for i in range(8): d[i << 27:(i + 1) << 27].copy_(h[i << 27:(i + 1) << 27], non_blocking=True)
    tk = None
    if "k" in order:
        dev.wait(0); tk = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize(); return tk, (time.perf_counter() - t0) * 1e3
for _ in range(2): run(["k", "c"])
for order in (["k"], ["c"], ["k", "c"], ["c", "k"], ["c8", "k"]):
    r = [run(order) for _ in range(3)]
    print(order, "kernel done ms", min(x[0] for x in r) if r[0][0] else None, "all done ms", min(x[1] for x in r))
