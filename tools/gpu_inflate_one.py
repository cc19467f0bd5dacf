"""Inflate a handful of zlib streams with one kernel variant (debug helper for compute-sanitizer)."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G  # noqa: E402
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nchunks = int(sys.argv[2]) if len(sys.argv) > 2 else 40
capi.lib().bitar_tune_inflate_variant(variant)
seg = 59460
mode = sys.argv[3] if len(sys.argv) > 3 else "lineitem"
levels = [(1, 0)]
if mode == "edge":
    cases = synth.edge_cases(seg)
    cases["lineitem"] = synth.lineitem_like(4 * seg)
    chunks = []
    for name, d in cases.items():
        for off in range(0, max(d.size, 1), seg):
            chunks.append(d[off:off + seg])
    levels = [(0, 0), (1, 0), (6, 0), (9, 0), (1, zlib.Z_FIXED)]
else:
    data = synth.lineitem_like(nchunks * seg)
    chunks = [data[i * seg:(i + 1) * seg] for i in range(nchunks)]
comps, origs = [], []
for c in chunks:
    for lvl, strat in levels:
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
        comps.append(np.frombuffer(co.compress(c.tobytes()) + co.flush(), np.uint8).copy())
        origs.append(c)
chunks = origs
pre = os.environ.get("PRE_VARIANT")
if pre is not None:
    capi.lib().bitar_tune_inflate_variant(int(pre))
    dev = G.open_device(seg)
    outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(c.size, 1) for c in chunks], src_shift=0, dst_shift=0)
    print("pre", pre, err)
    dev.close()
    capi.lib().bitar_tune_inflate_variant(variant)
dev = G.open_device(seg)
outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(c.size, 1) for c in chunks], src_shift=1, dst_shift=5)
import ctypes as C
if hasattr(capi.lib(), "bitar_debug_lane"):
    dbg = (C.c_uint * 16)()
    capi.lib().bitar_debug_lane(dbg)
    print("dbg", [hex(x) for x in dbg])
print("err", err, "status", set(res["status"].tolist()))
bad = [i for i, (o, c) in enumerate(zip(outs, chunks)) if not np.array_equal(o, c)]
print("n", len(chunks), "bad", bad[:20])
dev.close()
