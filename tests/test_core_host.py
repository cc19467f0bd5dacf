"""CPU: the kernels' host/device-shared source (bitar_b200/csrc/*.h) compiled for the host.

inflate_core.h runs with G = 1 (one lane per chunk): same table construction, header parsing, bit
reader, output ring/flush and error paths as on the GPU.  deflate_model.h is the sequential statement
of the deflate kernel.  Both are checked against the oracle; bit-exact."""
import zlib

import numpy as np
import pytest

import model_lib as M
import oracle_lib as O
from bitar_b200 import synth

SEG = 59460


def _chunks():
    cases = synth.edge_cases(SEG)
    cases["lineitem"] = synth.lineitem_like(3 * SEG)
    for name, d in cases.items():
        for off in range(0, max(d.size, 1), SEG):
            yield name, d[off:off + SEG]


def test_symbol_maps():
    assert M.lib().model_selfcheck() == 0


@pytest.mark.parametrize("lbits", [9, 10])
def test_inflate_core_vs_zlib(lbits):
    for name, ch in _chunks():
        for lvl, strat in [(0, 0), (1, 0), (9, 0), (1, zlib.Z_FIXED)]:
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
            z = np.frombuffer(co.compress(ch.tobytes()) + co.flush(), np.uint8)
            for mis in (0, 5):
                out, info = M.host_inflate(z, max(ch.size, 1), lbits, mis)
                assert info["status"] == 0 and info["guard_ok"] and info["consumed"] == z.size, (name, lvl, info)
                assert np.array_equal(out, ch), (name, lvl)


def test_inflate_core_errors():
    text = synth.edge_cases(SEG)["text"][:5000]
    z = np.frombuffer(zlib.compress(text.tobytes(), 6)[2:-4], np.uint8).copy()
    assert M.host_inflate(z[:100], 5000)[1]["status"] == 3          # truncated
    assert M.host_inflate(z, 4000)[1]["status"] == 1                # out of space
    bad = z.copy()
    bad[0] |= 6
    assert M.host_inflate(bad, 5000)[1]["status"] == 2              # BTYPE 3
    # over-subscribed dynamic header: BTYPE=2 with all 19 code-length codes of length 1
    over = np.frombuffer(bytes([0b00000101, 0xE0, 0xFF]) + bytes([0x49, 0x92, 0x24] * 3) + bytes(16), np.uint8)
    assert M.host_inflate(over, 100)[1]["status"] == 2
    # distance beyond the start of the output: fixed block, match first (len 3, dist 1)
    bits = "1" + "10" + "0000001" + "00000" + "0000000"   # BFINAL,BTYPE=1(01 lsb first), len sym 257, dist 0, EOB
    val = int(bits[::-1], 2)
    far = np.frombuffer(val.to_bytes(4, "little"), np.uint8)
    assert M.host_inflate(far, 100)[1]["status"] == 2
    with pytest.raises(ValueError):
        O.rfc_inflate(far, 100)


@pytest.mark.parametrize("huffman", [1, 2])
def test_deflate_model_streams_are_valid(huffman):
    total_model = total_zlib = 0
    for name, ch in _chunks():
        m = M.model_deflate(ch, huffman)
        assert np.array_equal(O.inflate_chunk(m, max(ch.size, 1)), ch), name
        out, info = O.rfc_inflate(m, max(ch.size, 1))
        assert np.array_equal(out, ch) and info["consumed"] == m.size
        out2, info2 = M.host_inflate(m, max(ch.size, 1))
        assert info2["status"] == 0 and np.array_equal(out2, ch)
        if name == "lineitem":
            total_model += m.size
            total_zlib += O.deflate_chunk(ch, 1, 15, huffman).size
    assert total_model <= 1.05 * total_zlib       # north_star ratio tolerance, on the columnar workload


def test_deflate_model_large_chunks_multi_block():
    data = synth.lineitem_like(3 * 65536 + 1000)
    m = M.model_deflate(data, 2)
    out, info = O.rfc_inflate(m, data.size)
    assert np.array_equal(out, data) and info["blocks"] == 4
    rnd = np.frombuffer(np.random.default_rng(5).bytes(65536 + 10), np.uint8)
    m = M.model_deflate(rnd, 2)
    assert np.array_equal(O.inflate_chunk(m, rnd.size), rnd)
    assert m.size <= rnd.size + 5 * 3              # never worse than three stored pieces (65535 + 1 + 10 bytes)
