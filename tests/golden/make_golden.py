"""Generates tests/golden/*.npz -- golden vectors for the hot path.

The reference has no fixtures (its test/ directory is an empty scaffold), so these are produced HERE
from the reference's codec dependency itself: CPython's zlib module (zlib 1.3 runtime, the library
DPDK's compress_zlib PMD calls) with the parameters bitar resolves to (src/config.cc:83-105):
raw deflate, level 1, window 15, memLevel 8, Z_DEFAULT_STRATEGY | Z_FIXED.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bitar_b200 import synth  # noqa: E402

SEG = 59460
OUT = os.path.dirname(os.path.abspath(__file__))


def zraw(b, level, strategy):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(b) + co.flush()


def main():
    # 1. per-chunk facts for seeded inputs (inputs are regenerated from the seed in the tests)
    rows = []
    cases = synth.edge_cases(SEG)
    cases["lineitem"] = synth.lineitem_like(3 * SEG)
    for name, d in sorted(cases.items()):
        for off in range(0, max(d.size, 1), SEG):
            ch = d[off:off + SEG].tobytes()
            dyn = zraw(ch, 1, zlib.Z_DEFAULT_STRATEGY)
            fix = zraw(ch, 1, zlib.Z_FIXED)
            rows.append((name, off, len(ch), len(dyn), hashlib.sha256(dyn).hexdigest(), len(fix),
                         hashlib.sha256(fix).hexdigest(), zlib.crc32(ch), zlib.adler32(ch),
                         hashlib.sha256(ch).hexdigest()))
    np.savez_compressed(
        os.path.join(OUT, "chunk_facts.npz"),
        name=np.array([r[0] for r in rows]), offset=np.array([r[1] for r in rows], np.int64),
        size=np.array([r[2] for r in rows], np.int64), dyn_len=np.array([r[3] for r in rows], np.int64),
        dyn_sha=np.array([r[4] for r in rows]), fix_len=np.array([r[5] for r in rows], np.int64),
        fix_sha=np.array([r[6] for r in rows]), crc32=np.array([r[7] for r in rows], np.uint32),
        adler32=np.array([r[8] for r in rows], np.uint32), data_sha=np.array([r[9] for r in rows]))
    # 2. a handful of real reference streams (small) for inflate tests: every block type
    text = cases["text"][:6000].tobytes()
    streams = {
        "text_l1_dynamic": (zraw(text, 1, zlib.Z_DEFAULT_STRATEGY), text),
        "text_l9_dynamic": (zraw(text, 9, zlib.Z_DEFAULT_STRATEGY), text),
        "text_l1_fixed": (zraw(text, 1, zlib.Z_FIXED), text),
        "text_l0_stored": (zraw(text, 0, zlib.Z_DEFAULT_STRATEGY), text),
        "lineitem_l1": (zraw(cases["lineitem"][:8192].tobytes(), 1, zlib.Z_DEFAULT_STRATEGY), cases["lineitem"][:8192].tobytes()),
        "zeros_l1": (zraw(bytes(SEG), 1, zlib.Z_DEFAULT_STRATEGY), bytes(SEG)),
        "empty_l1": (zraw(b"", 1, zlib.Z_DEFAULT_STRATEGY), b""),
    }
    np.savez_compressed(os.path.join(OUT, "ref_streams.npz"),
                        **{k + "__comp": np.frombuffer(v[0], np.uint8) for k, v in streams.items()},
                        **{k + "__plain": np.frombuffer(v[1], np.uint8) for k, v in streams.items()})
    print("zlib runtime", zlib.ZLIB_RUNTIME_VERSION, "rows", len(rows))


if __name__ == "__main__":
    main()
