"""Inflate speed on FOREIGN streams (zlib-produced, no parallel-inflate index).
usage: python tools/gpu_foreign_inflate.py [MiB] [level] [variants...]
variant 0 = default dispatch (speculative kernel, whole-stream kernel for what it declines), 6 = whole-stream kernel for
every stream without an index, 5 = whole-stream kernel only; 0:T = default dispatch with T output bytes per speculative range"""
import os
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
variants = sys.argv[3:] or ["0", "6"]
seg = 59460
data = synth.lineitem_like(mib << 20)
n = (data.size + seg - 1) // seg


def comp(i):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8)
    return c.compress(data[i * seg:(i + 1) * seg].tobytes()) + c.flush()


with ThreadPoolExecutor(16) as ex:
    streams = list(ex.map(comp, range(n)))
offs = np.zeros(n, np.uint64)
sizes = np.array([len(s) for s in streams], np.uint32)
offs[1:] = np.cumsum((sizes[:-1].astype(np.uint64) + 15) // 16 * 16)
buf = np.zeros(int(offs[-1]) + int(sizes[-1]) + 64, np.uint8)
for o, s in zip(offs, streams):
    buf[int(o):int(o) + len(s)] = np.frombuffer(s, np.uint8)
dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg))
src = torch.from_numpy(buf).cuda()
ref = torch.from_numpy(data).cuda()
out = torch.zeros(n * seg + 64, dtype=torch.uint8, device="cuda")
ops = np.zeros(n, capi.CHUNK_DTYPE)
ops["src"] = np.uint64(src.data_ptr()) + offs
ops["src_len"] = sizes
ops["dst"] = np.uint64(out.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(seg)
ops["dst_cap"] = seg
print(f"zlib level {level}: ratio {data.size / sizes.sum():.3f}, {n} streams")
for vs in variants:
    v, _, tgt = vs.partition(":")
    v = int(v)
    capi.lib().bitar_tune_inflate_variant(v)
    capi.lib().bitar_tune_spec_target(int(tgt or 0))
    out.zero_()
    best = 1e9
    for _ in range(3):
        dev.enqueue("inflate", 0, ops)
        dev.wait(0)
        k, t = dev.last_ms(0)
        best = min(best, k)
    cnt = np.zeros(8, np.uint32)
    capi.lib().bitar_debug_inflate_counters(dev._h, 0, cnt.ctypes.data)
    print(f"variant {vs}: no index {cnt[1]}, declined {cnt[7]}; {best:.3f} ms {data.size / best / 1e6:.1f} GB/s ok={bool((out[:data.size] == ref).all().item())}", flush=True)
dev.close()
