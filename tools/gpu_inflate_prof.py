"""Runs the inflate kernels of one or more variants (comma separated) a few times on GPU-compressed lineitem data
(for ncu, and for A/B runs).  usage: python tools/gpu_inflate_prof.py [variant[,variant...]] [MiB] [seg]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402

variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
seg = int(sys.argv[3]) if len(sys.argv) > 3 else 59460
data = synth.lineitem_like(mib << 20)
n = (data.size + seg - 1) // seg
dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 8))
src = torch.from_numpy(data).cuda()
out = torch.zeros(n * seg + 64, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ops, slots = dev.compress_ops(src.data_ptr(), data.size)
res = dev.enqueue("deflate", 0, ops)
dev.wait(0)
iops = dev.decompress_ops(slots, res["produced"], out.data_ptr())
for variant in variants:
    capi.lib().bitar_tune_inflate_variant(variant)
    out.zero_()
    best = 1e9
    for _ in range(4):
        dev.enqueue("inflate", 0, iops)
        dev.wait(0)
        k, t = dev.last_ms(0)
        best = min(best, k)
    print(f"variant {variant} {mib} MiB seg {seg}: kernels {best:.3f} ms (best of 4), call {t:.3f} ms, {data.size / best / 1e6:.1f} GB/s ok={bool((out[:data.size] == src).all().item())}", flush=True)
capi.lib().bitar_tune_inflate_variant(0)
dev.close()
