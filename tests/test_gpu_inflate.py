"""GPU parity: the sm_100a inflate kernel vs the oracle (zlib through oracle/, RFC 1951 restatement).

Reference-compressed streams (zlib raw deflate, the compress_zlib PMD's codec) must inflate on the GPU
to the original bytes -- BASELINE.json config 5 -- for every block type zlib can emit.
Bit-exact comparison (byte data)."""
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


def corpus(seg=SEG):
    cases = synth.edge_cases(seg)
    cases["lineitem"] = synth.lineitem_like(4 * seg)
    chunks = []
    for name, d in cases.items():
        for off in range(0, max(d.size, 1), seg):
            chunks.append((name, d[off:off + seg]))
    return chunks


@pytest.mark.parametrize("variant", [0, 5])
def test_reference_streams_inflate_on_gpu(cuda_device, variant):
    capi.lib().bitar_tune_inflate_variant(variant)
    dev = G.open_device(SEG)
    try:
        chunks = corpus()
        comps, origs = [], []
        for name, ch in chunks:
            for lvl, strat in [(0, 0), (1, 0), (6, 0), (9, 0), (1, zlib.Z_FIXED)]:
                comps.append(zraw(ch, lvl, strat))
                origs.append(ch)
        for shift in (0, 1):
            outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(o.size, 1) for o in origs],
                                                  src_shift=shift, dst_shift=4 * shift + shift)
            assert err is None, err
            assert (res["status"] == 0).all()
            for o, ref in zip(outs, origs):
                assert o.size == ref.size and np.array_equal(o, ref)
    finally:
        dev.close()
        capi.lib().bitar_tune_inflate_variant(0)


@pytest.mark.parametrize("variant", [0])
@pytest.mark.parametrize("huffman", [capi.HUFFMAN_DYNAMIC, capi.HUFFMAN_FIXED])
def test_gpu_streams_inflate_through_the_index(cuda_device, variant, huffman):
    """Chunks produced by the deflate kernel carry the parallel-inflate index: the sub-range kernel must
    give back the original bytes (bit-exact), for every edge case, alignment and multi-block chunk; chunks
    without an index (stored, tiny) take the whole-stream kernel in the same call."""
    capi.lib().bitar_tune_inflate_variant(variant)
    seg = 3 * 65536 + 5000
    dev = G.open_device(seg, huffman_enc=huffman)
    try:
        chunks = [c for _, c in corpus()]
        chunks += [synth.lineitem_like(seg), np.concatenate([np.frombuffer(np.random.default_rng(9).bytes(65536), np.uint8),
                                                             synth.lineitem_like(30000)])]
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        import model_lib as M
        n_indexed = sum(M.split_index(c)[1] is not None for c in comps)
        assert n_indexed >= 20
        for shift in (0, 3):
            outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(c.size, 1) for c in chunks], src_shift=shift, dst_shift=5 * shift)
            assert err is None, err
            assert (res["status"] == 0).all()
            for o, ref in zip(outs, chunks):
                assert o.size == ref.size and np.array_equal(o, ref)
        # damaged index entries and a short output buffer are reported, never silently wrong
        big = [i for i, c in enumerate(comps) if M.split_index(c)[1] is not None][:3]
        bad = [comps[i].copy() for i in big]
        bad[0][bad[0].size - 12 - 4 * 3] ^= 0x10         # a sub_bit entry
        bad[1][bad[1].size - 12 - 4 * 7 + 1] ^= 0x01
        outs, res, err = G.gpu_inflate_chunks(dev, bad, [chunks[big[0]].size, chunks[big[1]].size, chunks[big[2]].size - 1])
        assert err is not None and err.code == capi.E_IO_ERROR
        assert list(res["status"]) == [capi.OP_DATA_ERROR, capi.OP_DATA_ERROR, capi.OP_OUT_OF_SPACE]
    finally:
        dev.close()
        capi.lib().bitar_tune_inflate_variant(0)


@pytest.mark.parametrize("seg", [2049, 4096, 5000, 16384])
def test_small_segments_four_blocks_per_warp(cuda_device, seg):
    """Segments of at most 8 sub-ranges take the GROUP = 8 kernel (four blocks per warp, tables per group):
    round trip, checksums, a tiny last segment, single-sub-range and stored segments in the same call, and
    damaged chunks reported per op."""
    data = np.concatenate([synth.lineitem_like(40 * seg + 77), np.frombuffer(np.random.default_rng(4).bytes(2 * seg), np.uint8),
                           np.zeros(3 * seg + 5, np.uint8)])
    chunks = [data[o:o + seg] for o in range(0, data.size, seg)]
    dev = G.open_device(seg, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        for shift in (0, 5):
            outs, ires, err = G.gpu_inflate_chunks(dev, comps, [seg] * len(chunks), src_shift=shift % 4, dst_shift=shift)
            assert err is None and (ires["status"] == 0).all()
            for c, o, r in zip(chunks, outs, ires):
                assert np.array_equal(o, c)
                assert int(r["checksum"]) & 0xFFFFFFFF == O.crc32(c) and int(r["checksum"]) >> 32 == O.adler32(c)
        bad = [c.copy() for c in comps[:8]]
        for k, b in enumerate(bad):
            b[b.size - 13 - 4 * (k % 3)] ^= 0x20          # index words of the first chunks
        outs, ires, err = G.gpu_inflate_chunks(dev, bad + comps[8:12], [seg] * 12)
        assert err is not None and (ires["status"][:8] != 0).all() and (ires["status"][8:] == 0).all()
    finally:
        dev.close()


def test_corrupted_streams_fuzz(cuda_device):
    """600 randomly damaged chunks (GPU streams with their index, zlib streams) in one call: every op reports a
    status, nothing is written outside the destinations (guard bytes), the queue pair stays usable."""
    rng = np.random.default_rng(5)
    ch = synth.lineitem_like(SEG)
    dev = G.open_device(SEG)
    try:
        comps, _, _ = G.gpu_deflate_chunks(dev, [ch])
        good = [comps[0], zraw(ch, 1)]
        bad = []
        for t in range(600):
            s = good[t & 1].copy()
            for _ in range(int(rng.integers(1, 4))):
                k = s.size - 1 - int(rng.integers(0, 140)) if (t % 6 == 0) else int(rng.integers(0, s.size))
                s[k] ^= 1 << int(rng.integers(0, 8))
            bad.append(s)
        outs, res, err = G.gpu_inflate_chunks(dev, bad, [SEG] * len(bad))     # asserts the guard bytes
        assert set(int(x) for x in res["status"]) <= {0, 1, 2, 3}
        assert (res["produced"] <= SEG).all()
        outs, res, err = G.gpu_inflate_chunks(dev, good, [SEG, SEG])
        assert err is None and all(np.array_equal(o, ch) for o in outs)
    finally:
        dev.close()


def test_host_resident_buffers_are_staged(cuda_device):
    """Compressed slots and the destination in PINNED HOST memory: the inflate call gathers / scatters through
    device memory (capi.cu stage_copy_kernel); results and guard bytes as for device buffers."""
    import ctypes as C
    L = capi.lib()
    data = synth.lineitem_like(37 * SEG + 4321)
    n = (data.size + SEG - 1) // SEG
    dev = G.open_device(SEG, slot_mem_kind=capi.MEM_PINNED, max_preallocate_memzones=n + 8)
    h_in, h_out = C.c_void_p(), C.c_void_p()
    try:
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, data.size, 64, C.byref(h_in)))
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, n * SEG + 64, 64, C.byref(h_out)))
        C.memmove(h_in.value, data.ctypes.data, data.size)
        C.memset(h_out.value, 0xA5, n * SEG + 64)
        ops, slots = dev.compress_ops(h_in.value, data.size)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        for dst_off in (0, 3):                                  # aligned and misaligned destination
            iops = dev.decompress_ops(slots, res["produced"], h_out.value + dst_off)
            ires = dev.enqueue("inflate", 0, iops)
            dev.wait(0)
            back = np.ctypeslib.as_array(C.cast(h_out.value, C.POINTER(C.c_uint8)), shape=(n * SEG + 64,))
            assert int(ires["produced"].sum()) == data.size and (ires["status"] == 0).all()
            assert np.array_equal(back[dst_off:dst_off + data.size], data)
            assert (back[dst_off + data.size:] == 0xA5).all(), "inflate wrote past the produced bytes"
            C.memset(h_out.value, 0xA5, n * SEG + 64)
        # zlib-produced streams in pinned memory (no index: the speculative kernel) go through the same staging
        zs, zp = O.compress_buffer(data, SEG)
        h_z = C.c_void_p()
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, zs.size, 64, C.byref(h_z)))
        C.memmove(h_z.value, zs.ctypes.data, zs.size)
        ptrs = np.uint64(h_z.value) + np.arange(n, dtype=np.uint64) * np.uint64(zs.shape[1])
        ires = dev.enqueue("inflate", 0, dev.decompress_ops(ptrs, zp, h_out.value))
        dev.wait(0)
        back = np.ctypeslib.as_array(C.cast(h_out.value, C.POINTER(C.c_uint8)), shape=(n * SEG + 64,))
        assert int(ires["produced"].sum()) == data.size and np.array_equal(back[:data.size], data)
        capi.check(L.bitar_mem_free(capi.MEM_PINNED, 0, h_z))
        assert sum(dev.put_slot(s) for s in slots[::-1]) == n
    finally:
        for b in (h_in, h_out):
            if b.value:
                L.bitar_mem_free(capi.MEM_PINNED, 0, b)
        dev.close()


@pytest.mark.parametrize("ck", [capi.CHECKSUM_NONE, capi.CHECKSUM_CRC32_ADLER32])
def test_staged_calls_in_batches_on_lane_streams(cuda_device, ck):
    """A staged inflate call large enough for several batches runs them on the queue pair's lane streams, each batch
    with its own slice of the task / counter / checksum buffers.  Forced here on a small buffer (1 MiB per batch):
    contiguous destination (copy-engine copy-back), scattered destination (scatter kernel), foreign streams mixed in
    (no index: the speculative kernel, the whole-stream kernel for what it declines), checksums, two queue pairs at once."""
    import ctypes as C
    L = capi.lib()
    L.bitar_tune_stage_batch.argtypes = [C.c_ulonglong]
    data = synth.lineitem_like(160 * SEG + 777)
    n = (data.size + SEG - 1) // SEG
    dev = G.open_device(SEG, qps=2, slot_mem_kind=capi.MEM_PINNED, max_preallocate_memzones=n + 8, checksum_type=ck)
    bufs = []

    def pinned(nbytes):
        p = C.c_void_p()
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, nbytes, 64, C.byref(p)))
        bufs.append(p)
        return p.value

    try:
        L.bitar_tune_stage_batch(1 << 20)
        h_in, h_out = pinned(data.size), pinned(n * SEG + 64)
        C.memmove(h_in, data.ctypes.data, data.size)
        ops, slots = dev.compress_ops(h_in, data.size)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        # every fifth chunk replaced by a zlib-produced stream (no index -> whole-stream kernel inside the batches)
        zs, zp = O.compress_buffer(data, SEG)
        h_z = pinned(zs.size)
        C.memmove(h_z, zs.ctypes.data, zs.size)
        src = np.array(slots, dtype=np.uint64)
        produced = res["produced"].copy()
        for i in range(0, n, 5):
            src[i] = h_z + i * zs.shape[1]
            produced[i] = zp[i]
        want_ck = [(O.adler32(data[i * SEG:(i + 1) * SEG]) << 32) | O.crc32(data[i * SEG:(i + 1) * SEG]) for i in range(n)]
        back = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint8)), shape=(n * SEG + 64,))
        # (a) contiguous destination, both queue pairs at once on halves of the buffer
        C.memset(h_out, 0xA5, n * SEG + 64)
        half = n // 2
        r0 = dev.enqueue("inflate", 0, dev.decompress_ops(src[:half], produced[:half], h_out))
        r1 = dev.enqueue("inflate", 1, dev.decompress_ops(src[half:], produced[half:], h_out + half * SEG))
        dev.wait(0)
        dev.wait(1)
        ires = np.concatenate([r0, r1])
        assert (ires["status"] == 0).all() and int(ires["produced"].sum()) == data.size
        assert np.array_equal(back[:data.size], data) and (back[data.size:] == 0xA5).all()
        if ck:
            assert [int(c) for c in ires["checksum"]] == want_ck
        # (b) scattered destination: segment i at a stride of SEG + 48 (not the Decompress() layout)
        stride = SEG + 48
        h_sc = pinned(n * stride + 64)
        C.memset(h_sc, 0x5A, n * stride + 64)
        iops = dev.decompress_ops(src, produced, h_sc)
        iops["dst"] = np.uint64(h_sc) + np.arange(n, dtype=np.uint64) * np.uint64(stride) + np.uint64(1)
        ires = dev.enqueue("inflate", 0, iops)
        dev.wait(0)
        sc = np.ctypeslib.as_array(C.cast(h_sc, C.POINTER(C.c_uint8)), shape=(n * stride + 64,))
        assert (ires["status"] == 0).all()
        for i in range(n):
            c = data[i * SEG:(i + 1) * SEG]
            assert np.array_equal(sc[i * stride + 1:i * stride + 1 + c.size], c), i
            assert (sc[i * stride + 1 + c.size:(i + 1) * stride + 1] == 0x5A).all(), i
        if ck:
            assert [int(c) for c in ires["checksum"]] == want_ck
        # (c) every source a pool slot (constant stride): the batches gather with strided copy-engine transfers
        C.memset(h_out, 0xA5, n * SEG + 64)
        ires = dev.enqueue("inflate", 1, dev.decompress_ops(np.array(slots, dtype=np.uint64), res["produced"], h_out + 5))
        dev.wait(1)
        assert (ires["status"] == 0).all() and int(ires["produced"].sum()) == data.size
        assert np.array_equal(back[5:5 + data.size], data) and (back[5 + data.size:] == 0xA5).all()
        if ck:
            assert [int(c) for c in ires["checksum"]] == want_ck
    finally:
        L.bitar_tune_stage_batch(0)
        for b in bufs:
            L.bitar_mem_free(capi.MEM_PINNED, 0, b)
        dev.close()


def test_bitar_decompress_contract(cuda_device):
    """Decompress(): op i lands at out + i*S, total = sum(produced) (src/device.cc:240-318); checked
    against oracle_decompress_buffer on the oracle's own compressed slots."""
    import torch
    data = synth.lineitem_like(40 * SEG + 12345)
    slots, produced = O.compress_buffer(data, SEG)
    ref, _ = O.decompress_buffer(slots, produced, SEG)
    assert np.array_equal(ref, data)
    dev = G.open_device(SEG)
    try:
        d_slots = G.to_dev(slots.reshape(-1))
        out = torch.zeros(slots.shape[0] * SEG, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ptrs = np.uint64(d_slots.data_ptr()) + np.arange(slots.shape[0], dtype=np.uint64) * np.uint64(slots.shape[1])
        ops = dev.decompress_ops(ptrs, produced, out.data_ptr())
        res = dev.enqueue("inflate", 0, ops)
        dev.wait(0)
        total = int(res["produced"].sum())
        assert total == data.size
        assert np.array_equal(out[:total].cpu().numpy(), data)
    finally:
        dev.close()


def test_inflate_error_statuses(cuda_device):
    dev = G.open_device(SEG)
    try:
        text = synth.edge_cases(SEG)["text"][:20000]
        good = zraw(text, 6)
        bad_btype = good.copy()
        bad_btype[0] |= 6                      # BTYPE = 3
        stored = zraw(np.frombuffer(np.random.default_rng(3).bytes(1000), np.uint8), 0)
        bad_nlen = stored.copy()
        bad_nlen[3] ^= 0xFF                    # LEN != ~NLEN
        comps = [good, good[:good.size // 2], good, bad_btype, bad_nlen, good]
        caps = [text.size, text.size, 1000, text.size, 1000, text.size]
        outs, res, err = G.gpu_inflate_chunks(dev, comps, caps)
        assert err is not None and err.code == capi.E_IO_ERROR      # src/device.cc:512-520
        assert list(res["status"]) == [capi.OP_OK, capi.OP_TRUNCATED, capi.OP_OUT_OF_SPACE,
                                       capi.OP_DATA_ERROR, capi.OP_DATA_ERROR, capi.OP_OK]
        assert np.array_equal(outs[0], text) and np.array_equal(outs[5], text)
        # the queue pair is usable again after a failed call (the reference leaves it busy: quirk not copied)
        outs, res, err = G.gpu_inflate_chunks(dev, [good], [text.size])
        assert err is None and np.array_equal(outs[0], text)
    finally:
        dev.close()


@pytest.mark.parametrize("ck", [capi.CHECKSUM_CRC32_ADLER32, capi.CHECKSUM_CRC32, capi.CHECKSUM_ADLER32])
def test_inflate_checksums(cuda_device, ck):
    """Checksums of the inflated bytes == zlib's, for zlib streams (whole-stream kernel) and for the deflate
    kernel's own streams (sub-range kernel: per-lane partial sums combined per block and per chunk), including
    multi-block chunks and a chunk that mixes a stored and a coded block."""
    dev = G.open_device(3 * 65536 + 5000, checksum_type=ck)
    try:
        chunks = [c for _, c in corpus()][:40]
        chunks += [synth.lineitem_like(3 * 65536 + 5000), synth.lineitem_like(65536 + 1),
                   np.concatenate([np.frombuffer(np.random.default_rng(9).bytes(65536), np.uint8), synth.lineitem_like(30000)])]
        gpu_comps, _, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        for comps in ([zraw(c, 1) for c in chunks], gpu_comps):
            outs, res, err = G.gpu_inflate_chunks(dev, comps, [max(c.size, 1) for c in chunks])
            assert err is None
            for c, o, r in zip(chunks, outs, res):
                assert np.array_equal(o, c)
                want_crc = O.crc32(c) if ck & capi.CHECKSUM_CRC32 else 0
                want_adler = O.adler32(c) if ck & capi.CHECKSUM_ADLER32 else 0
                assert int(r["checksum"]) & 0xFFFFFFFF == want_crc, c.size
                assert int(r["checksum"]) >> 32 == want_adler, c.size
    finally:
        dev.close()
