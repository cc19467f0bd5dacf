"""GPU parity: streams WITHOUT a parallel-inflate index (zlib-produced: what the reference's own CPU path writes,
/root/reference/src/memory.cc:432-505 takes any buffer) through the speculative lane-parallel kernel
(bitar_b200/csrc/inflate_spec_kernel.cuh).  Bit-exact against the input and against the whole-stream kernel; what the
speculative kernel declines must come out of the whole-stream kernel with that kernel's status words."""
import zlib

import numpy as np
import pytest

import gpu_util as G
import oracle_lib as O
from bitar_b200 import _capi as capi
from bitar_b200 import synth

pytestmark = pytest.mark.gpu
SEG = 59460


def zraw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return np.frombuffer(co.compress(data.tobytes()) + co.flush(), np.uint8).copy()


def counters(dev, qp=0):
    out = np.zeros(8, np.uint32)
    capi.check(capi.lib().bitar_debug_inflate_counters(dev._h, qp, out.ctypes.data))
    return {"tasks": int(out[0]), "no_index": int(out[1]), "declined": int(out[7])}


def test_benchmark_streams_take_the_speculative_kernel(cuda_device):
    """zlib level-1 streams of BASELINE config 2's chunks: every one is decoded by the speculative kernel (none
    declined), bit-exact, with zlib's checksums."""
    data = synth.lineitem_like(150 * SEG + 1234)
    chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
    comps = [zraw(c, 1) for c in chunks]
    dev = G.open_device(SEG, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        for shift in (0, 3):
            outs, res, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks], src_shift=shift, dst_shift=7 * shift)
            assert err is None, err
            c = counters(dev)
            assert c["no_index"] == len(comps) and c["declined"] == 0, c
            for i, (o, ref) in enumerate(zip(outs, chunks)):
                assert np.array_equal(o, ref), i
                assert int(res["checksum"][i]) & 0xFFFFFFFF == O.crc32(ref) and int(res["checksum"][i]) >> 32 == O.adler32(ref)
    finally:
        dev.close()


@pytest.mark.parametrize("seg", [4096, SEG, 1 << 18, 1 << 20])
def test_speculative_and_whole_stream_kernels_agree(cuda_device, seg):
    """Every block type zlib emits, the ratio corpus and the edge cases, at 4 KiB .. 1 MiB segments: the default path
    (speculative kernel, whole-stream kernel for what it declines) and the whole-stream kernel alone (variant 6) return
    the same bytes, produced counts and status words."""
    inputs = dict(synth.ratio_corpus(max(seg, 1 << 16)))
    inputs.update(synth.edge_cases(min(seg, SEG)))
    comps, origs = [], []
    for name, d in inputs.items():
        ch = d[:seg]
        for lvl, strat in [(1, 0), (6, 0), (9, 0), (1, zlib.Z_FIXED), (1, zlib.Z_HUFFMAN_ONLY), (1, zlib.Z_RLE), (0, 0)]:
            comps.append(zraw(ch, lvl, strat))
            origs.append(ch)
    dev = G.open_device(seg)
    try:
        got = {}
        for variant in (0, 6):
            capi.lib().bitar_tune_inflate_variant(variant)
            for target in ((0, 200, 1536) if variant == 0 else (0,)):
                capi.lib().bitar_tune_spec_target(target)
                outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(o.size, 1) for o in origs], src_shift=variant // 6, dst_shift=3)
                assert err is None, err
                assert (res["status"] == 0).all()
                for o, ref in zip(outs, origs):
                    assert o.size == ref.size and np.array_equal(o, ref)
                got[(variant, target)] = counters(dev)
        # (level 0, incompressible and tiny inputs are stored blocks or too short for a round: declined; the rest is not)
        assert got[(0, 0)]["no_index"] == len(comps) and got[(0, 0)]["declined"] < len(comps), got
        assert got[(6, 0)]["declined"] == 0, got
    finally:
        capi.lib().bitar_tune_inflate_variant(0)
        capi.lib().bitar_tune_spec_target(0)
        dev.close()


def test_damaged_streams_same_verdict_as_the_whole_stream_kernel(cuda_device):
    """Bit flips and truncation: with and without the speculative kernel every op reports the same status and produced
    count (a declined stream is decoded by the whole-stream kernel), a stream that still decodes gives the same bytes,
    and nothing is written outside the destinations."""
    rng = np.random.default_rng(31)
    ch = synth.lineitem_like(SEG)
    good = [zraw(ch, 1), zraw(ch, 6), zraw(synth.text_source(SEG), 1)]
    bad = []
    for t in range(400):
        s = good[t % 3].copy()
        if t % 5 == 0:
            s = s[:int(rng.integers(8, s.size))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, s.size))] ^= 1 << int(rng.integers(0, 8))
        bad.append(s)
    caps = [SEG if t % 7 else SEG // 2 for t in range(len(bad))]
    # streams whose only fault is ONE match that reaches below the start of the output, in the first range or in a later,
    # speculatively decoded one (a lane cannot check a distance against an output position it does not know: phase B does)
    from test_core_host import _fixed_stream
    n_crafted = 0
    for where in (0, 3000, 9000, 20000):
        tokens, out = [], bytearray()
        bad_at = None
        while len(out) < 40000:
            if bad_at is None and len(out) >= where:
                bad_at = len(tokens)
            if len(out) > 300 and rng.random() < 0.35:
                ln, dist = int(rng.integers(3, 40)), int(rng.integers(1, min(len(out), 32768) + 1))
                tokens.append(("M", ln, dist))
                for _ in range(ln):
                    out.append(out[-dist])
            else:
                b = int(rng.integers(0, 256))
                tokens.append(("L", b))
                out.append(b)
        pos = sum(1 if t[0] == "L" else t[1] for t in tokens[:bad_at])
        for extra in (1, 700):
            bad.append(_fixed_stream(tokens[:bad_at] + [("M", 5, pos + extra)] + tokens[bad_at:]))
            caps.append(SEG)
            n_crafted += 1
    dev = G.open_device(SEG)
    try:
        runs = []
        for variant in (0, 6):
            capi.lib().bitar_tune_inflate_variant(variant)
            outs, res, err = G.gpu_inflate_chunks(dev, bad, caps)      # asserts the guard bytes
            runs.append((outs, res["status"].copy(), res["produced"].copy()))
        assert np.array_equal(runs[0][1], runs[1][1])
        assert (runs[0][1][-n_crafted:] == capi.OP_DATA_ERROR).all()
        ok = runs[0][1] == 0
        assert np.array_equal(runs[0][2][ok], runs[1][2][ok])
        for i in np.nonzero(ok)[0]:
            assert np.array_equal(runs[0][0][i], runs[1][0][i]), i
    finally:
        capi.lib().bitar_tune_inflate_variant(0)
        dev.close()
