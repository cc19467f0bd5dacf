"""One pass of every kernel over the edge-case corpus, small enough for compute-sanitizer:
deflate (dynamic + fixed, with checksums), indexed inflate, whole-stream inflate of zlib streams, staged
inflate from pinned host memory.   usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
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

SEG = 59460


def main():
    cases = synth.edge_cases(SEG)
    chunks = []
    for name in ("random_7", "random_59460", "zeros_59461", "ab", "period258", "text", "lowentropy", "random_0"):
        d = cases[name]
        chunks += [d[o:o + SEG] for o in range(0, max(d.size, 1), SEG)][:2]
    chunks += [synth.lineitem_like(2 * 65536 + 777), synth.lineitem_like(SEG), synth.lineitem_like(5000)]
    for huff, ck in ((capi.HUFFMAN_DYNAMIC, capi.CHECKSUM_CRC32_ADLER32), (capi.HUFFMAN_FIXED, capi.CHECKSUM_NONE)):
        dev = G.open_device(2 * 65536 + 777, huffman_enc=huff, checksum_type=ck)
        comps, res, err = G.gpu_deflate_chunks(dev, chunks, src_shift=1, dst_shift=3)
        assert err is None
        outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(c.size, 1) for c in chunks], src_shift=1, dst_shift=5)
        assert err is None and all(np.array_equal(o, c) for o, c in zip(outs, chunks))
        zs = []
        for c in chunks:
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            zs.append(np.frombuffer(co.compress(c.tobytes()) + co.flush(), np.uint8).copy())
        outs, res, err = G.gpu_inflate_chunks(dev, zs, [max(c.size, 1) for c in chunks])
        assert err is None and all(np.array_equal(o, c) for o, c in zip(outs, chunks))
        # damaged streams: error paths must stay in bounds too
        bad = [z.copy() for z in comps[-3:]]
        for b in bad:
            b[b.size // 2] ^= 0x55
            b[-14] ^= 0x03
        G.gpu_inflate_chunks(dev, bad, [chunks[-3].size, chunks[-2].size, chunks[-1].size])
        dev.close()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
