"""GPU parity: the sm_100a deflate kernel.

 * every GPU stream must inflate under the reference's zlib path (oracle) AND the independent RFC 1951
   restatement to the original bytes (bit-exact);
 * the stream must be bit-identical to the sequential model of the kernel (tools/model), which pins
   the match finder, Huffman construction and bit packing;
 * checksums must equal zlib's;
 * compressed size within 5 % of zlib level 1 on the columnar workload (north_star tolerance)."""
import numpy as np
import pytest

import gpu_util as G
import model_lib as M
import oracle_lib as O
from bitar_b200 import _capi as capi
from bitar_b200 import synth

pytestmark = pytest.mark.gpu
SEG = 59460
RATIO_TOLERANCE = 1.05   # GPU bytes <= 1.05 x zlib level-1 bytes


def corpus(seg=SEG):
    cases = synth.edge_cases(seg)
    cases["lineitem"] = synth.lineitem_like(6 * seg)
    chunks = []
    for name, d in cases.items():
        for off in range(0, max(d.size, 1), seg):
            chunks.append(d[off:off + seg])
    return chunks


@pytest.mark.parametrize("huffman", [capi.HUFFMAN_DYNAMIC, capi.HUFFMAN_FIXED])
def test_gpu_streams_inflate_under_zlib_and_match_model(cuda_device, huffman):
    dev = G.open_device(SEG, huffman_enc=huffman, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        chunks = corpus()
        for shift in (0, 3):
            outs, res, err = G.gpu_deflate_chunks(dev, chunks, src_shift=shift, dst_shift=shift)
            assert err is None, err
            assert (res["status"] == 0).all()
            for c, comp, r in zip(chunks, outs, res):
                back = O.inflate_chunk(comp, max(c.size, 1))          # zlib, the reference codec
                assert np.array_equal(back, c)
                back2, info = O.rfc_inflate(comp, max(c.size, 1))     # independent decoder
                body, index = M.split_index(comp)                     # the parallel-inflate index follows the stream
                assert np.array_equal(back2, c) and info["consumed"] == body.size
                model = M.model_deflate(c, huffman)
                assert np.array_equal(comp, model), "GPU stream differs from the kernel model"
                assert int(r["checksum"]) & 0xFFFFFFFF == O.crc32(c)
                assert int(r["checksum"]) >> 32 == O.adler32(c)
    finally:
        dev.close()


def test_known_answers(cuda_device):
    """random chunk -> stored blocks; empty -> 03 00; zeros compress > 200x."""
    dev = G.open_device(SEG)
    try:
        rnd = np.frombuffer(np.random.default_rng(1).bytes(SEG), np.uint8)
        outs, res, err = G.gpu_deflate_chunks(dev, [rnd, np.zeros(0, np.uint8), np.zeros(SEG, np.uint8)])
        assert err is None
        assert outs[0].size == SEG + 5 and outs[0][0] == 0x01          # one stored block, BFINAL
        assert outs[1].tobytes() == b"\x03\x00"
        assert outs[2].size < SEG // 200
    finally:
        dev.close()


def test_out_of_space(cuda_device):
    dev = G.open_device(SEG)
    try:
        rnd = np.frombuffer(np.random.default_rng(2).bytes(5000), np.uint8)
        outs, res, err = G.gpu_deflate_chunks(dev, [rnd, rnd], cap=4096)
        assert err is not None and err.code == capi.E_IO_ERROR
        assert (res["status"] == capi.OP_OUT_OF_SPACE).all() and (res["produced"] == 0).all()
    finally:
        dev.close()


@pytest.mark.parametrize("seg", [4096, 16384, 59460, 65536, 262144, 1048576])
def test_ratio_and_roundtrip_by_segment_size(cuda_device, seg):
    """Compress() then Decompress() == identity, ratio next to zlib -1 on identical chunks."""
    import torch
    data = synth.lineitem_like(24 * (1 << 20) if seg >= 65536 else 6 * (1 << 20))
    dev = G.open_device(seg, max_preallocate_memzones=max(20, data.size // seg + 8))
    try:
        src = G.to_dev(data)
        ops, slots = dev.compress_ops(src.data_ptr(), data.size)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        gpu_bytes = int(res["produced"].sum())
        zslots, zprod = O.compress_buffer(data, seg, threads=8)
        assert gpu_bytes <= RATIO_TOLERANCE * int(zprod.sum()), (gpu_bytes, int(zprod.sum()))
        out = torch.zeros(len(ops) * seg + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        iops = dev.decompress_ops(slots, res["produced"], out.data_ptr())
        ires = dev.enqueue("inflate", 0, iops)
        dev.wait(0)
        assert int(ires["produced"].sum()) == data.size
        assert np.array_equal(out[:data.size].cpu().numpy(), data)
        # spot-check a few GPU streams under zlib
        host = {}
        for i in (0, len(ops) // 2, len(ops) - 1):
            n = int(res["produced"][i])
            t = torch.empty(n, dtype=torch.uint8, device="cuda")
            capi.check(capi.lib().bitar_qp_memcpy(dev._h, 0, t.data_ptr(), int(slots[i]), n))
            dev.wait(0)
            chunk = data[i * seg:(i + 1) * seg]
            assert np.array_equal(O.inflate_chunk(t.cpu().numpy(), chunk.size), chunk)
        assert sum(dev.put_slot(s) for s in slots[::-1]) == len(slots)
    finally:
        dev.close()


def test_random_structured_ragged_batch(cuda_device):
    """One call, 300 ops of ragged sizes (1 .. 200 000 bytes: below a sub-range, across sub-range and 64 KiB block
    boundaries) and random structure: every GPU stream equals the sequential model, inflates under zlib, and the
    GPU inflates the whole batch back (sub-range kernel and whole-stream kernel mixed in one call)."""
    from test_core_host import _structured
    rng = np.random.default_rng(77)
    sizes = [1, 2, 3, 2047, 2048, 2049, 4095, 4097, 65535, 65536, 65537, 131072, 200000] + \
            [int(rng.integers(1, 70000)) for _ in range(287)]
    chunks = [_structured(rng, n) for n in sizes]
    dev = G.open_device(200000, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        comps, res, err = G.gpu_deflate_chunks(dev, chunks, src_shift=1, dst_shift=2)
        assert err is None and (res["status"] == 0).all()
        for i, (c, z, r) in enumerate(zip(chunks, comps, res)):
            assert np.array_equal(z, M.model_deflate(c, capi.HUFFMAN_DYNAMIC)), (i, c.size)
            if i % 5 == 0:
                assert np.array_equal(O.inflate_chunk(z, c.size), c)
            assert int(r["checksum"]) & 0xFFFFFFFF == O.crc32(c) and int(r["checksum"]) >> 32 == O.adler32(c)
        outs, ires, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks], src_shift=3, dst_shift=1)
        assert err is None
        for c, o, r, r0 in zip(chunks, outs, ires, res):
            assert np.array_equal(o, c) and int(r["checksum"]) == int(r0["checksum"])
    finally:
        dev.close()


@pytest.mark.parametrize("huffman", [capi.HUFFMAN_DYNAMIC, capi.HUFFMAN_FIXED])
def test_small_chunk_instance_matches_model(cuda_device, huffman):
    """Calls whose chunks all fit 16 KiB run on the 4-warp instance of the deflate kernel (bitar::dks): same bytes as
    the sequential model, valid under zlib, checksums equal zlib's; empty, 1-byte, incompressible and all-zero chunks
    and every sub-range boundary included."""
    from test_core_host import _structured
    rng = np.random.default_rng(99)
    sizes = [0, 1, 2, 3, 7, 8, 2047, 2048, 2049, 4095, 4096, 4097, 8192, 12288, 16383, 16384] + \
            [int(rng.integers(1, 16385)) for _ in range(240)]
    chunks = [_structured(rng, n) for n in sizes]
    chunks += [np.frombuffer(rng.bytes(n), np.uint8) for n in (100, 4096, 16384)]          # stored
    chunks += [np.zeros(n, np.uint8) for n in (4096, 16384)]                               # long matches
    chunks += [synth.lineitem_like(3 * 16384)[i * 16384:(i + 1) * 16384] for i in range(3)]
    dev = G.open_device(16384, huffman_enc=huffman, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        comps, res, err = G.gpu_deflate_chunks(dev, chunks, src_shift=5, dst_shift=9)
        assert err is None and (res["status"] == 0).all()
        for i, (c, z, r) in enumerate(zip(chunks, comps, res)):
            assert np.array_equal(z, M.model_deflate(c, huffman)), (i, c.size)
            assert np.array_equal(O.inflate_chunk(z, max(c.size, 1)), c)
            assert int(r["checksum"]) & 0xFFFFFFFF == O.crc32(c) and int(r["checksum"]) >> 32 == O.adler32(c)
        outs, ires, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks])
        assert err is None
        for c, o in zip(chunks, outs):
            assert np.array_equal(o, c)
    finally:
        dev.close()
