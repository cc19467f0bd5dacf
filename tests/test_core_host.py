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


def test_huffman_codes_are_complete():
    """Length-limited code construction (shared with the deflate kernel): complete (Kraft sum 1) and within
    15 / 7 bits for skewed frequency sets whose optimal trees are far deeper than the limit."""
    assert M.lib().model_huffman_fuzz(60000, 7) == 0


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


@pytest.mark.parametrize("lbits", [8, 9, 10])
def test_inflate_fast_lane_vs_zlib(lbits):
    """inflate_fast.h (the production lane-per-stream decoder) with one lane on the CPU: every block type
    zlib emits, both alignments, checksums folded in on the way out."""
    for name, ch in _chunks():
        for lvl, strat in [(0, 0), (1, 0), (6, 0), (9, 0), (1, zlib.Z_FIXED)]:
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
            z = np.frombuffer(co.compress(ch.tobytes()) + co.flush(), np.uint8)
            for mis in (0, 5):
                out, info = M.host_inflate_fast(z, max(ch.size, 1), lbits, mis, checksum_type=3)
                assert info["status"] == 0 and info["guard_ok"] and info["consumed"] == z.size, (name, lvl, info)
                assert np.array_equal(out, ch), (name, lvl)
                assert info["crc32"] == O.crc32(ch) and info["adler32"] == O.adler32(ch), (name, lvl)


def test_inflate_fast_lane_errors():
    text = synth.edge_cases(SEG)["text"][:5000]
    z = np.frombuffer(zlib.compress(text.tobytes(), 6)[2:-4], np.uint8).copy()
    for lbits in (8, 9, 10):
        assert M.host_inflate_fast(z[:100], 5000, lbits)[1]["status"] == 3          # truncated
        out, info = M.host_inflate_fast(z, 4000, lbits)
        assert info["status"] == 1 and info["guard_ok"] and info["produced"] <= 4000  # out of space
        assert np.array_equal(out, text[:info["produced"]])
        bad = z.copy()
        bad[0] |= 6
        assert M.host_inflate_fast(bad, 5000, lbits)[1]["status"] == 2              # BTYPE 3
        over = np.frombuffer(bytes([0b00000101, 0xE0, 0xFF]) + bytes([0x49, 0x92, 0x24] * 3) + bytes(16), np.uint8)
        assert M.host_inflate_fast(over, 100, lbits)[1]["status"] == 2              # over-subscribed code-length code
        bits = "1" + "10" + "0000001" + "00000" + "0000000"
        far = np.frombuffer(int(bits[::-1], 2).to_bytes(4, "little"), np.uint8)
        assert M.host_inflate_fast(far, 100, lbits)[1]["status"] == 2               # distance before the start
        stored = np.frombuffer(zlib.compressobj(0, zlib.DEFLATED, -15).compress(bytes(1000)) +
                               zlib.compressobj(0, zlib.DEFLATED, -15).flush(), np.uint8)
    rnd = np.frombuffer(np.random.default_rng(3).bytes(1000), np.uint8)
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    st = np.frombuffer(co.compress(rnd.tobytes()) + co.flush(), np.uint8).copy()
    st[3] ^= 0xFF
    assert M.host_inflate_fast(st, 1000)[1]["status"] == 2                          # LEN != ~NLEN
    # tiny capacity, every length: the deferred capacity check never writes past dst_cap
    for cap in (1, 2, 15, 16, 17, 31, 33, 100):
        out, info = M.host_inflate_fast(z, cap)
        assert info["status"] == 1 and info["guard_ok"] and np.array_equal(out, text[:info["produced"]]), cap


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
        body, index = M.split_index(m)
        assert np.array_equal(out, ch) and info["consumed"] == body.size
        # the parallel-inflate index: present iff the chunk has more than one sub-range and a coded block
        assert (index is not None) == (ch.size > M.SUB and (body[0] & 6) != 0), name
        if index is not None:
            assert index["total_out"] == ch.size
        out2, info2 = M.host_inflate(m, max(ch.size, 1))
        assert info2["status"] == 0 and np.array_equal(out2, ch)
        if name == "lineitem":
            total_model += m.size
            total_zlib += O.deflate_chunk(ch, 1, 15, huffman).size
    # north_star ratio tolerance (5 % of zlib level 1, dynamic Huffman) on the columnar workload; the fixed
    # code pays more for the matches the 2 KiB sub-range rule and the 1024-entry tables give up (DESIGN.md "Ratio"): 10 % there
    assert total_model <= (1.05 if huffman == 2 else 1.10) * total_zlib


@pytest.mark.parametrize("huffman", [1, 2])
def test_indexed_inflate_of_model_streams(huffman):
    """Sub-range parallel decode through the index == the original bytes; a damaged index is detected."""
    seen = 0
    for name, ch in _chunks():
        m = M.model_deflate(ch, huffman)
        body, index = M.split_index(m)
        for mis in (0, 7):
            out, info = M.host_inflate_indexed(m, max(ch.size, 1), mis)
            assert info["indexed"] == (index is not None), name
            if index is None:
                continue
            assert info["status"] == 0 and info["guard_ok"] and np.array_equal(out, ch), (name, info)
            seen += 1
        if index is not None and name == "lineitem":
            bad = m.copy()
            k = m.size - 12 - 4 * 5      # a sub_bit entry
            bad[k] ^= 1
            assert M.host_inflate_indexed(bad, ch.size)[1]["status"] == 2
            assert M.host_inflate_indexed(m, ch.size - 1)[1]["status"] == 1      # total_out > capacity
    assert seen > 10
    big = synth.lineitem_like(3 * 65536 + 5000)
    m = M.model_deflate(big, huffman)
    out, info = M.host_inflate_indexed(m, big.size)
    assert info["status"] == 0 and info["subs"] == 3 * 32 + 3 and np.array_equal(out, big)
    mixed = np.concatenate([np.frombuffer(np.random.default_rng(9).bytes(65536), np.uint8), synth.lineitem_like(30000)])
    m = M.model_deflate(mixed, huffman)       # a stored block followed by a coded one
    out, info = M.host_inflate_indexed(m, mixed.size)
    assert info["indexed"] and info["status"] == 0 and np.array_equal(out, mixed)


def test_corrupted_streams_never_write_out_of_bounds():
    """Random damage to the stream or to the index: every decoder (whole-stream lane, indexed sub-range path)
    must report something and keep its writes inside [dst, dst + cap) -- DEFLATE carries no integrity check, so
    a damaged stream may also decode 'successfully' to other bytes; what must never happen is a hang or a write
    outside the destination."""
    rng = np.random.default_rng(11)
    ch = synth.lineitem_like(SEG)
    m = M.model_deflate(ch, 2)
    z = np.frombuffer(zlib.compress(ch.tobytes(), 1)[2:-4], np.uint8)
    for trial in range(300):
        for stream, fn in ((m, M.host_inflate_indexed), (m, M.host_inflate_fast), (z, M.host_inflate_fast)):
            bad = stream.copy()
            for _ in range(int(rng.integers(1, 4))):
                if trial % 3 == 0 and stream is m:
                    k = stream.size - 1 - int(rng.integers(0, 140))      # inside the index
                else:
                    k = int(rng.integers(0, stream.size))
                bad[k] ^= 1 << int(rng.integers(0, 8))
            out, info = fn(bad, SEG)
            assert info["guard_ok"] and info["status"] in (0, 1, 2, 3), (trial, info)
            if fn is M.host_inflate_indexed and info["indexed"] and info["status"] == 0:
                assert out.size == SEG


def test_gzip_member_framing_of_model_streams():
    """engine.gzip_members(): kernel-format chunks (index stripped) + CRC-32 / ISIZE per member form a
    multi-member gzip file that Python's gzip module decompresses to the original buffer."""
    import gzip
    from bitar_b200 import engine as E
    data = synth.lineitem_like(3 * SEG + 100)
    chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
    comps = [M.model_deflate(c, 2) for c in chunks]
    res = np.zeros(len(chunks), dtype=[("produced", "<u4"), ("status", "<u4"), ("checksum", "<u8")])
    for i, c in enumerate(chunks):
        res["checksum"][i] = O.crc32(c)
        assert E.stream_length(comps[i]) == M.split_index(comps[i])[0].size
    assert gzip.decompress(E.gzip_members(comps, res, [c.size for c in chunks])) == data.tobytes()


def test_unframe_locates_the_raw_stream():
    """engine.unframe(): offsets / lengths / trailers of zlib streams and gzip members (with optional header fields)."""
    import gzip
    import io
    from bitar_b200 import engine as E
    data = synth.lineitem_like(20000).tobytes()
    z = np.frombuffer(zlib.compress(data, 6), np.uint8)
    off, ln, kind, exp = E.unframe(z)
    assert (off, kind) == (2, "zlib") and exp == (zlib.adler32(data),)
    assert zlib.decompress(z[off:off + ln].tobytes(), -15) == data
    bio = io.BytesIO()
    with gzip.GzipFile(filename="some-name.bin", mode="wb", fileobj=bio, compresslevel=1, mtime=7) as f:   # FNAME set
        f.write(data)
    g = np.frombuffer(bio.getvalue(), np.uint8)
    off, ln, kind, exp = E.unframe(g)
    assert kind == "gzip" and exp == (zlib.crc32(data), len(data)) and off == 10 + len("some-name.bin") + 1
    assert zlib.decompress(g[off:off + ln].tobytes(), -15) == data
    for bad in (b"\x78\x20" + bytes(10), b"\x1f\x8b\x07" + bytes(20), b"hello world, not framed"):
        with pytest.raises(ValueError):
            E.unframe(np.frombuffer(bad, np.uint8))


def test_zlib_stream_framing_of_model_streams():
    """engine.zlib_streams(): 78 01 + stream (index stripped) + Adler-32 is what zlib.decompress() accepts
    (RFC 1950; it verifies the Adler-32)."""
    import zlib
    from bitar_b200 import engine as E
    data = synth.lineitem_like(2 * SEG + 77)
    chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
    comps = [M.model_deflate(c, 2) for c in chunks]
    res = np.zeros(len(chunks), dtype=[("produced", "<u4"), ("status", "<u4"), ("checksum", "<u8")])
    for i, c in enumerate(chunks):
        res["checksum"][i] = np.uint64(O.adler32(c)) << np.uint64(32)
    for c, z in zip(chunks, E.zlib_streams(comps, res)):
        assert zlib.decompress(z) == c.tobytes()
    bad = bytearray(E.zlib_streams(comps, res)[0])
    bad[-1] ^= 1
    with pytest.raises(zlib.error):
        zlib.decompress(bytes(bad))


def test_deflate_model_large_chunks_multi_block():
    data = synth.lineitem_like(3 * 65536 + 1000)
    m = M.model_deflate(data, 2)
    out, info = O.rfc_inflate(m, data.size)
    body, index = M.split_index(m)
    assert np.array_equal(out, data) and info["blocks"] == 4 and info["consumed"] == body.size
    assert index is not None and [b["len"] for b in index["blocks"]] == [65536, 65536, 65536, 1000]
    assert all(len(b["sub_bit"]) == (b["len"] + 2047) // 2048 for b in index["blocks"])
    rnd = np.frombuffer(np.random.default_rng(5).bytes(65536 + 10), np.uint8)
    m = M.model_deflate(rnd, 2)
    assert np.array_equal(O.inflate_chunk(m, rnd.size), rnd)
    assert m.size <= rnd.size + 5 * 3              # never worse than three stored pieces (65535 + 1 + 10 bytes)


def _structured(rng, n):
    """Random bytes with random structure: runs, periodic patterns, small alphabets, copies from far back."""
    out = np.empty(n, np.uint8)
    at = 0
    while at < n:
        kind = int(rng.integers(0, 6))
        ln = int(min(n - at, rng.integers(1, 700)))
        if kind == 0:
            out[at:at + ln] = rng.integers(0, 256, ln, dtype=np.uint8)
        elif kind == 1:
            out[at:at + ln] = int(rng.integers(0, 256))
        elif kind == 2:
            per = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)
            out[at:at + ln] = np.resize(per, ln)
        elif kind == 3:
            out[at:at + ln] = rng.integers(0, int(rng.integers(2, 17)), ln, dtype=np.uint8)
        elif at > 0:
            back = int(rng.integers(1, min(at, 40000) + 1))
            for k in range(ln):            # overlapping copy semantics
                out[at + k] = out[at + k - back]
        else:
            out[at:at + ln] = 7
        at += ln
    return out


def test_random_structured_inputs_round_trip_on_the_cpu_builds():
    """200 random inputs of ragged sizes (1 .. 70 000 bytes, crossing the 2 KiB sub-range and 64 KiB block
    boundaries): model stream -> zlib, -> independent RFC 1951 decoder, -> whole-stream lane, -> indexed path."""
    rng = np.random.default_rng(2026)
    sizes = [1, 2, 3, 2047, 2048, 2049, 4095, 4097, 65535, 65536, 65537] + [int(rng.integers(1, 70000)) for _ in range(189)]
    for i, n in enumerate(sizes):
        d = _structured(rng, n)
        for huffman in ((2,) if i % 4 else (1, 2)):
            m = M.model_deflate(d, huffman)
            assert np.array_equal(O.inflate_chunk(m, n), d), (i, n)
            body, index = M.split_index(m)
            out, info = O.rfc_inflate(m, n)
            assert np.array_equal(out, d) and info["consumed"] == body.size
            out, info = M.host_inflate_fast(m, n, 9, i % 7)
            assert info["status"] == 0 and info["guard_ok"] and np.array_equal(out, d)
            out, info = M.host_inflate_indexed(m, n, i % 5)
            assert info["indexed"] == (index is not None)
            if index is not None:
                assert info["status"] == 0 and info["guard_ok"] and np.array_equal(out, d), (i, n, info)


@pytest.mark.parametrize("name", ["lineitem_mix", "sorted_int64", "dict_int32", "price_f64", "text_source", "elf_binary",
                                  "char10_strings", "word_strings", "period4096_rows"])
def test_model_ratio_corpus_within_tolerance_of_zlib_level_1(name):
    """The sequential model of the deflate kernel (the GPU tests pin the kernel to it bit for bit) on every input of the
    ratio corpus: at most 5 % more bytes than zlib level 1 on identical 59 460-byte chunks, every stream valid under
    zlib and under the two-phase decoder of the inflate kernels (host build)."""
    from bitar_b200 import synth
    data = synth.ratio_corpus(768 << 10)[name]
    seg = 59460
    model_bytes = zlib_bytes = 0
    for off in range(0, data.size, seg):
        ch = data[off:off + seg]
        z = M.model_deflate(ch, 2)
        model_bytes += z.size
        zlib_bytes += O.deflate_chunk(ch, 1, 15, O.HUFFMAN_DYNAMIC).size
        assert np.array_equal(O.inflate_chunk(z, ch.size), ch)
        if off == 0:
            out, info = M.host_inflate_indexed(z, ch.size)
            assert info["status"] == 0 and info["indexed"] and np.array_equal(out, ch)
    assert model_bytes <= 1.05 * zlib_bytes, (name, model_bytes, zlib_bytes, model_bytes / zlib_bytes)


def _zraw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return np.frombuffer(co.compress(data.tobytes()) + co.flush(), np.uint8).copy()


@pytest.mark.parametrize("level,strategy", [(1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (1, zlib.Z_FIXED), (1, zlib.Z_HUFFMAN_ONLY), (1, zlib.Z_RLE), (0, zlib.Z_DEFAULT_STRATEGY)])
def test_speculative_lanes_decode_streams_without_an_index(level, strategy):
    """inflate_spec.h (the lane-parallel decoder of zlib-produced streams: 32 ranges per round start at GUESSED bit
    offsets and join the true symbol chain) on the CPU: whatever it does not decline is bit-exact, for every block type,
    alignment and size; stored blocks are declined (the whole-stream kernel's business)."""
    inputs = dict(synth.ratio_corpus(1 << 18))
    inputs.update({k: v for k, v in synth.edge_cases(SEG).items() if v.size in (SEG, SEG + 1) or k in ("ab", "period258", "period32768")})
    decoded = 0
    for name, d in inputs.items():
        for seg in (4096, SEG, 1 << 18):
            ch = d[:seg]
            if ch.size == 0:
                continue
            comp = _zraw(ch, level, strategy)
            for mis, target in ((0, 768), (5, 300), (0, 1536)):
                out, info = M.host_inflate_spec(comp, ch.size, target, mis)
                assert info["guard_ok"], (name, seg, info)
                if level == 0:
                    assert info["declined"], (name, seg, info)
                    continue
                if not info["declined"]:
                    assert np.array_equal(out, ch), (name, seg, level, strategy, info)
                    decoded += 1
    assert level == 0 or decoded > 50


def test_speculative_lanes_keep_most_lanes_busy_on_the_benchmark_streams():
    """zlib level-1 streams of BASELINE config 2's chunks: nothing is declined, a round uses most of its 32 lanes and
    nearly every lane finds its successor within the recorded steps (the walk is a few symbols long)."""
    data = synth.lineitem_like(24 * SEG)
    tot = {"rounds": 0, "ranges": 0, "short_rounds": 0, "walk_symbols": 0, "blocks": 0}
    for i in range(24):
        ch = data[i * SEG:(i + 1) * SEG]
        out, info = M.host_inflate_spec(_zraw(ch, 1), SEG, 768)
        assert not info["declined"] and np.array_equal(out, ch) and info["guard_ok"]
        for k in tot:
            tot[k] += info[k]
    assert tot["ranges"] >= 20 * tot["rounds"], tot
    assert tot["walk_symbols"] <= 16 * tot["ranges"], tot


def test_speculative_lanes_on_damaged_streams_agree_with_zlib_or_decline():
    """DEFLATE carries no integrity check: a damaged stream may still be a valid stream.  The speculative decoder may
    decline anything, but what it does decode must be what zlib decodes from the same bytes -- and it never writes
    outside the destination."""
    rng = np.random.default_rng(23)
    ch = synth.lineitem_like(SEG)
    good = [_zraw(ch, 1), _zraw(ch, 6), _zraw(synth.text_source(SEG), 1)]
    kept = 0
    for trial in range(240):
        bad = good[trial % 3].copy()
        if trial % 5 == 0:
            bad = bad[:int(rng.integers(8, bad.size))]                        # truncated
        else:
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, bad.size))] ^= 1 << int(rng.integers(0, 8))
        out, info = M.host_inflate_spec(bad, SEG, 768, trial % 4)
        assert info["guard_ok"] and info["produced"] <= SEG, (trial, info)
        if info["declined"]:
            continue
        d = zlib.decompressobj(-15)
        try:
            ref = d.decompress(bad.tobytes(), SEG + 1)
        except zlib.error:
            ref = None
        assert ref is not None and d.eof and len(ref) <= SEG, (trial, info)
        assert np.array_equal(out, np.frombuffer(ref, np.uint8)), trial
        kept += 1
    assert kept > 0


def _fixed_stream(tokens):
    """A raw DEFLATE stream of one fixed-Huffman block from ('L', byte) / ('M', length, distance) tokens (RFC 1951 3.2.6)."""
    lb = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
    le = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
    db = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]
    de = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
    bits = []

    def put(v, n):                      # LSB first (header fields, extra bits)
        bits.extend((v >> i) & 1 for i in range(n))

    def code(v, n):                     # Huffman codes go MSB first
        bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))

    def litlen(sym):
        if sym < 144: code(0x30 + sym, 8)
        elif sym < 256: code(0x190 + sym - 144, 9)
        elif sym < 280: code(sym - 256, 7)
        else: code(0xC0 + sym - 280, 8)

    put(1, 1)
    put(1, 2)
    for t in tokens:
        if t[0] == "L":
            litlen(t[1])
        else:
            _, ln, dist = t
            s = max(i for i in range(29) if lb[i] <= ln and (i != 28 or ln == 258))
            if ln == 258:
                s = 28
            litlen(257 + s)
            put(ln - lb[s], le[s])
            d = max(i for i in range(30) if db[i] <= dist)
            code(d, 5)
            put(dist - db[d], de[d])
    litlen(256)
    bits.extend([0] * (-len(bits) % 8))
    return np.packbits(np.array(bits, np.uint8), bitorder="little")


def test_speculative_lanes_decline_a_match_that_reaches_below_the_output():
    """A lane of the speculative decoder does not know where its range lies in the output, so it cannot check a distance
    against it; phase B does (per byte in the kernel, sp::resolve_range_serial here).  A stream whose ONLY fault is one
    match that reaches below the start of the output -- in the first range or in a later, speculatively decoded one --
    must be declined; the same stream without it decodes."""
    rng = np.random.default_rng(41)
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
        good = _fixed_stream(tokens)
        ref = np.frombuffer(bytes(out), np.uint8)
        assert zlib.decompress(good.tobytes(), -15) == bytes(out)
        got, info = M.host_inflate_spec(good, ref.size, 768)
        assert not info["declined"] and np.array_equal(got, ref)
        # the faulty twin: a match at output position `pos` whose distance is pos + 1 .. (one byte too far, or more)
        pos = sum(1 if t[0] == "L" else t[1] for t in tokens[:bad_at])
        for extra in (1, 700):
            if pos + extra > 32768:
                continue
            faulty = tokens[:bad_at] + [("M", 5, pos + extra)] + tokens[bad_at:]
            bad = _fixed_stream(faulty)
            with pytest.raises(zlib.error):
                zlib.decompress(bad.tobytes(), -15)
            got, info = M.host_inflate_spec(bad, ref.size + 5, 768)
            assert info["declined"] and info["guard_ok"], (where, extra, info)


def test_speculative_lanes_on_random_structured_inputs():
    """Randomly structured buffers (runs, periods, small alphabets, far copies) of random sizes, compressed by zlib at a
    level and strategy per case: the speculative decoder either declines or returns the input, whatever range size it
    aims at and however the buffers are aligned."""
    rng = np.random.default_rng(77)
    decoded = 0
    for case in range(40):
        n = int(rng.integers(200, 200000))
        ch = _structured(rng, n)
        lvl, strat = [(1, 0), (6, 0), (9, 0), (1, zlib.Z_HUFFMAN_ONLY), (1, zlib.Z_RLE), (1, zlib.Z_FIXED)][case % 6]
        comp = _zraw(ch, lvl, strat)
        target = int(rng.integers(64, 1537))
        out, info = M.host_inflate_spec(comp, n, target, case % 7)
        assert info["guard_ok"], (case, info)
        if not info["declined"]:
            assert np.array_equal(out, ch), (case, n, lvl, strat, target, info)
            decoded += 1
        # a destination one byte short is never written past: declined, or produced <= capacity
        out, info = M.host_inflate_spec(comp, n - 1, target, case % 5)
        assert info["guard_ok"] and info["declined"], (case, info)
    assert decoded >= 20
