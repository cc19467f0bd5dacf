"""Does a compress that reads pinned host memory (SM bulk loads over PCIe) run beside a device-to-host copy?
Times each alone and both together."""
import ctypes as C
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

L = capi.lib()
seg, U = 59460, 1 << 30
n = (U + seg - 1) // seg
data = synth.lineitem_like(U)
dev = CompressDevice(0, 2).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 64, slot_mem_kind=capi.MEM_PINNED))
h_in, h_out = C.c_void_p(), C.c_void_p()
capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, U, 64, C.byref(h_in)))
capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, U, 64, C.byref(h_out)))
C.memmove(h_in.value, data.ctypes.data, U)
d_tmp = torch.empty(U, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ops, slots = dev.compress_ops(h_in.value, U)


def run(compress, copy):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if compress:
        dev.enqueue("deflate", 0, ops)
    if copy:
        capi.check(L.bitar_qp_memcpy(dev._h, 1, h_out.value, d_tmp.data_ptr(), U))
    tc = tk = None
    if compress:
        dev.wait(0)
        tc = time.perf_counter() - t0
    if copy:
        dev.wait(1)
        tk = time.perf_counter() - t0
    return tc, tk


for _ in range(2):
    run(True, True)
a = run(True, False)[0]
b = run(False, True)[1]
c = run(True, True)
print(f"compress from pinned host alone {a * 1e3:.1f} ms ({U / a / 1e9:.1f} GB/s); D2H copy alone {b * 1e3:.1f} ms ({U / b / 1e9:.1f} GB/s); "
      f"together: compress {c[0] * 1e3:.1f} ms, copy {c[1] * 1e3:.1f} ms")
