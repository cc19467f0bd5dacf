"""Per-phase cycle breakdown of the deflate kernel (thread-0 clock64 counters, debug build hook)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402

NAMES = ["load", "match", "sort", "merge", "tables+hdr", "encode", "finish", "lengths+codes", "cl_rle", "cl_tree+type", "far", "-", "-", "-", "-", "-"]


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    seg = int(sys.argv[2]) if len(sys.argv) > 2 else 59460
    L = capi.lib()
    L.bitar_debug_deflate_profile.argtypes = [C.c_int, C.c_void_p]
    for wname in ["lineitem"] + list(synth.COLUMNS):
        data = synth.lineitem_like(mib << 20) if wname == "lineitem" else synth.column(wname, (mib << 20) // 4)
        n = (data.size + seg - 1) // seg
        dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 8))
        src = torch.from_numpy(data).cuda()
        torch.cuda.synchronize()
        ops, slots = dev.compress_ops(src.data_ptr(), data.size)
        for _ in range(2):
            dev.enqueue("deflate", 0, ops)
            dev.wait(0)
        L.bitar_debug_deflate_profile(1, None)
        dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        k, _ = dev.last_ms(0)
        out = np.zeros(16, np.uint64)
        L.bitar_debug_deflate_profile(0, out.ctypes.data)
        tot = float(out[:11].sum())
        print(f"[{wname}] kernel {k:.3f} ms ({data.size / k / 1e6:.1f} GB/s); cycles/chunk:",
              " ".join(f"{nm}={int(v) // n}({100 * v / tot:.0f}%)" for nm, v in zip(NAMES, out) if v), flush=True)
        dev.close()


if __name__ == "__main__":
    main()
