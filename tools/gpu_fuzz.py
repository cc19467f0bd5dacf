"""Extended randomized parity run on a GPU box (not part of the test suite: minutes, many seeds).
Every GPU stream must equal the sequential model (tools/model) byte for byte, and inflate on the GPU to the input;
chunk sizes cover both deflate instances (<= 16 KiB: 4-warp instance) and multi-block streams.  The same chunks compressed
by zlib (a level and strategy per chunk: no index) must inflate on the GPU to the input too, with zlib's checksums, through
the speculative kernel and whatever it declines.
usage: python tools/gpu_fuzz.py [seeds] [first_seed]"""
import os
import sys

import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G  # noqa: E402
import model_lib as M  # noqa: E402
from test_core_host import _structured  # noqa: E402
from bitar_b200 import _capi as capi  # noqa: E402

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
bad = 0
for seed in range(first, first + seeds):
    rng = np.random.default_rng(seed)
    for name, hi, cap, n in (("small", 16384, 16384, 400), ("large", 150000, 150000, 160)):
        sizes = [int(rng.integers(1, hi + 1)) for _ in range(n)]
        chunks = [_structured(rng, s) for s in sizes]
        huff = capi.HUFFMAN_DYNAMIC if seed % 3 else capi.HUFFMAN_FIXED
        dev = G.open_device(cap, huffman_enc=huff, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
        try:
            comps, res, err = G.gpu_deflate_chunks(dev, chunks, src_shift=seed % 16, dst_shift=(seed * 7) % 16)
            assert err is None and (res["status"] == 0).all(), err
            for i, (c, z) in enumerate(zip(chunks, comps)):
                if not np.array_equal(z, M.model_deflate(c, huff)):
                    bad += 1
                    print(f"MISMATCH seed {seed} {name} chunk {i} size {c.size}", flush=True)
            outs, ires, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks], src_shift=seed % 4, dst_shift=seed % 16)
            assert err is None, err
            for i, (c, o, r, r0) in enumerate(zip(chunks, outs, ires, res)):
                if not np.array_equal(o, c) or int(r["checksum"]) != int(r0["checksum"]):
                    bad += 1
                    print(f"ROUNDTRIP seed {seed} {name} chunk {i} size {c.size}", flush=True)
            # zlib-produced streams of the same chunks
            zs = []
            for i, c in enumerate(chunks):
                lvl, strat = [(1, 0), (6, 0), (9, 0), (1, zlib.Z_HUFFMAN_ONLY), (1, zlib.Z_RLE), (1, zlib.Z_FIXED), (0, 0), (1, 0)][(i + seed) % 8]
                co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
                zs.append(np.frombuffer(co.compress(c.tobytes()) + co.flush(), np.uint8).copy())
            outs, zres, err = G.gpu_inflate_chunks(dev, zs, [c.size for c in chunks], src_shift=(seed + 1) % 4, dst_shift=(seed * 3) % 16)
            assert err is None, err
            cnt = np.zeros(8, np.uint32)
            capi.lib().bitar_debug_inflate_counters(dev._h, 0, cnt.ctypes.data)
            for i, (c, o, r, r0) in enumerate(zip(chunks, outs, zres, res)):
                if not np.array_equal(o, c) or int(r["checksum"]) != int(r0["checksum"]):
                    bad += 1
                    print(f"FOREIGN seed {seed} {name} chunk {i} size {c.size}", flush=True)
            print(f"  seed {seed} {name}: {cnt[1]} zlib streams, {cnt[7]} declined by the speculative kernel", flush=True)
        finally:
            dev.close()
    print(f"seed {seed} done, mismatches so far {bad}", flush=True)
print("FUZZ", "FAILED" if bad else "OK", bad)
sys.exit(1 if bad else 0)
