"""CPU: pins the oracle (oracle/liboracle.so) before anything trusts it.

The reference carries no golden vectors (test/CMakeLists.txt:23-26 is empty), so the oracle is pinned
against (a) fixtures generated from the reference's codec dependency -- zlib 1.3 through CPython --
by tests/golden/make_golden.py, (b) the known-answer facts of SURVEY.md 8(c), and (c) the standard
CRC-32 / Adler-32 check values.  Bit-exact everywhere."""
import hashlib
import os
import zlib

import numpy as np
import pytest

import oracle_lib as O
from bitar_b200 import synth

SEG = 59460
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    cases = synth.edge_cases(SEG)
    cases["lineitem"] = synth.lineitem_like(3 * SEG)
    return cases


def test_chunk_facts_match_golden():
    g = np.load(os.path.join(GOLD, "chunk_facts.npz"))
    cases = _cases()
    assert len(g["name"]) == 42
    for i in range(len(g["name"])):
        ch = cases[str(g["name"][i])][int(g["offset"][i]):int(g["offset"][i]) + SEG]
        assert ch.size == int(g["size"][i]) and hashlib.sha256(ch.tobytes()).hexdigest() == str(g["data_sha"][i])
        dyn = O.deflate_chunk(ch, 1, 15, O.HUFFMAN_DYNAMIC)
        fix = O.deflate_chunk(ch, 1, 15, O.HUFFMAN_FIXED)
        assert dyn.size == int(g["dyn_len"][i]) and hashlib.sha256(dyn.tobytes()).hexdigest() == str(g["dyn_sha"][i])
        assert fix.size == int(g["fix_len"][i]) and hashlib.sha256(fix.tobytes()).hexdigest() == str(g["fix_sha"][i])
        assert O.crc32(ch) == int(g["crc32"][i]) == O.rfc_crc32(ch)
        assert O.adler32(ch) == int(g["adler32"][i]) == O.rfc_adler32(ch)


def test_reference_streams_golden():
    g = np.load(os.path.join(GOLD, "ref_streams.npz"))
    names = sorted({k.split("__")[0] for k in g.files})
    assert len(names) == 7
    for n in names:
        comp, plain = g[n + "__comp"], g[n + "__plain"]
        assert np.array_equal(O.inflate_chunk(comp, max(plain.size, 1)), plain)
        out, info = O.rfc_inflate(comp, max(plain.size, 1))
        assert np.array_equal(out, plain) and info["consumed"] == comp.size


def test_known_answers_survey_8c():
    rnd = np.frombuffer(np.random.default_rng(1).bytes(SEG), np.uint8)
    c = O.deflate_chunk(rnd)
    assert c.size == 59480 and c[0] == 0x00                      # 4 stored blocks x 5 bytes
    z = np.zeros(SEG, np.uint8)
    assert O.deflate_chunk(z).size == 277
    assert O.deflate_chunk(z, huffman=O.HUFFMAN_FIXED).size == 581
    assert O.deflate_chunk(z)[0] & 7 == 5                        # BFINAL=1, BTYPE=2
    assert O.deflate_chunk(z, huffman=O.HUFFMAN_FIXED)[0] & 7 == 3
    assert O.deflate_chunk(np.zeros(0, np.uint8)).tobytes() == b"\x03\x00"
    assert O.crc32(b"123456789") == 0xCBF43926 == O.rfc_crc32(b"123456789")
    assert O.adler32(b"Wikipedia") == 0x11E60398 == O.rfc_adler32(b"Wikipedia")
    assert O.crc32(b"") == 0 and O.adler32(b"") == 1 and O.rfc_crc32(b"") == 0 and O.rfc_adler32(b"") == 1
    assert O.lib().oracle_zlib_version() == b"1.3"


def test_segment_size_formula():
    """Configuration::UpdateCompressedSegSize, src/config.cc:59-73, and the limits of config.h:41-48."""
    assert O.compressed_seg_size(59460) == 65406
    assert O.compressed_seg_size(2048) == 4096
    assert O.compressed_seg_size(4095) == 4096        # smaller than the stored bound: SURVEY 7.2
    assert O.compressed_seg_size(8) == 16
    assert O.compressed_seg_size(32768) == 36044
    assert int(O.lib().oracle_max_seg_size()) == 59460 and int(O.lib().oracle_min_seg_size()) == 8
    assert O.stored_bound(59460) == 59465 and O.stored_bound(65536) == 65546 and O.stored_bound(0) == 5


@pytest.mark.parametrize("threads", [1, 3])
def test_buffer_contract_roundtrip(threads):
    """Compress -> Decompress over the chunking contract (src/device.cc:156-318)."""
    data = synth.lineitem_like(7 * SEG + 4321)
    slots, produced = O.compress_buffer(data, SEG, threads=threads)
    assert slots.shape[0] == 8 and (produced > 0).all()
    for i in range(8):   # every slot is a complete raw stream of exactly its segment
        seg = data[i * SEG:(i + 1) * SEG]
        assert zlib.decompressobj(-15).decompress(slots[i, :produced[i]].tobytes()) == seg.tobytes()
    out, got = O.decompress_buffer(slots, produced, SEG, threads=threads)
    assert np.array_equal(out, data) and got[-1] == 4321 and (got[:-1] == SEG).all()


def test_independent_decoder_agrees_with_zlib_on_all_levels():
    for name, d in _cases().items():
        ch = d[:SEG]
        for lvl, strat in [(0, 0), (1, 0), (6, 0), (9, 0), (1, zlib.Z_FIXED)]:
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
            z = np.frombuffer(co.compress(ch.tobytes()) + co.flush(), np.uint8)
            out, info = O.rfc_inflate(z, max(ch.size, 1))
            assert np.array_equal(out, ch) and info["consumed"] == z.size, (name, lvl)


def test_independent_decoder_rejects_invalid():
    good = np.frombuffer(zlib.compress(b"hello hello hello hello", 6)[2:-4], np.uint8).copy()
    with pytest.raises(ValueError):
        O.rfc_inflate(good[:3], 100)
    bad = good.copy()
    bad[0] |= 6
    with pytest.raises(ValueError):
        O.rfc_inflate(bad, 100)
