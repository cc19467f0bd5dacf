"""GPU parity, round 2: the ratio corpus, the configured window, bare streams, oversized ops, mixed staging.

The reference forwards `window_size` into both xforms (/root/reference/src/config.cc:83-105) and takes any buffer as
Decompress() input (src/memory.cc:432-505); the compressed size is gated against zlib level 1 (the reference's level,
src/config.cc:87) on every input of synth.ratio_corpus(), not only on the benchmark mix."""
import zlib

import numpy as np
import pytest
import torch

import gpu_util as G
import model_lib as M
import oracle_lib as O
from bitar_b200 import _capi as capi
from bitar_b200 import engine as E
from bitar_b200 import synth

pytestmark = pytest.mark.gpu
SEG = 59460
RATIO_TOLERANCE = 1.05   # compressed bytes / zlib level-1 bytes on identical chunks (north_star: "within 5 %")


def _chunks(data, seg=SEG):
    return [data[o:o + seg] for o in range(0, data.size, seg)]


@pytest.mark.parametrize("name", ["lineitem_mix", "sorted_int64", "dict_int32", "price_f64", "text_source", "elf_binary",
                                  "char10_strings", "word_strings", "period4096_rows"])
def test_ratio_corpus_within_tolerance_of_zlib_level_1(cuda_device, name):
    """Every corpus input: GPU bytes <= 1.05 x zlib -1 bytes on the same chunks, GPU stream == sequential model,
    zlib inflates the GPU streams, the GPU inflates them back."""
    data = synth.ratio_corpus(2 << 20)[name]
    chunks = _chunks(data)
    dev = G.open_device(SEG)
    try:
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None and (res["status"] == 0).all()
        gpu_bytes = sum(c.size for c in comps)
        zlib_bytes = sum(O.deflate_chunk(c, 1, 15, O.HUFFMAN_DYNAMIC).size for c in chunks)
        assert gpu_bytes <= RATIO_TOLERANCE * zlib_bytes, (name, gpu_bytes, zlib_bytes, gpu_bytes / zlib_bytes)
        for i in (0, len(chunks) // 2, len(chunks) - 1):
            assert np.array_equal(comps[i], M.model_deflate(chunks[i], capi.HUFFMAN_DYNAMIC)), (name, i)
            assert np.array_equal(O.inflate_chunk(comps[i], chunks[i].size), chunks[i])
        outs, ires, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks])
        assert err is None and (ires["status"] == 0).all()
        for c, o in zip(chunks, outs):
            assert np.array_equal(o, c)
    finally:
        dev.close()


@pytest.mark.parametrize("window", [8, 9, 10, 11, 12, 13, 14, 15])
def test_configured_window_is_honoured(cuda_device, window):
    """No match reaches farther back than 1 << window_size: zlib opened with the same window inflates every chunk
    (inflateInit2(-w) rejects "invalid distance too far back" otherwise)."""
    rng = np.random.default_rng(window)
    data = np.concatenate([synth.ratio_corpus(192 << 10)[k] for k in ("text_source", "period4096_rows", "lineitem_mix", "word_strings")])
    chunks = _chunks(data) + [np.tile(np.frombuffer(rng.bytes(700), np.uint8), 80)]
    dev = G.open_device(SEG, window_size=window)
    try:
        assert int(dev.cfg.window_size) == window
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None and (res["status"] == 0).all()
        for c, z in zip(chunks, comps):
            d = zlib.decompressobj(-window)
            out = d.decompress(E.strip_index(z).tobytes()) + d.flush()
            assert out == c.tobytes()
        outs, ires, err = G.gpu_inflate_chunks(dev, comps, [c.size for c in chunks])
        assert err is None
        for c, o in zip(chunks, outs):
            assert np.array_equal(o, c)
    finally:
        dev.close()


def test_bare_streams_without_the_index(cuda_device):
    """emit_index = False: `produced` is exactly the RFC 1951 stream (a consumer that checks consumed == length is
    satisfied), the streams equal the indexed ones minus the trailer, and they inflate here through the whole-stream
    kernel."""
    data = synth.lineitem_like(6 * SEG + 999)
    chunks = _chunks(data)
    dev0 = G.open_device(SEG)
    dev1 = G.open_device(SEG, emit_index=False)
    try:
        with_idx, _, err = G.gpu_deflate_chunks(dev0, chunks)
        assert err is None
        bare, res, err = G.gpu_deflate_chunks(dev1, chunks)
        assert err is None and (res["status"] == 0).all()
        for c, b, w in zip(chunks, bare, with_idx):
            assert M.split_index(b)[1] is None and (M.split_index(w)[1] is not None) == (c.size > 2048)
            assert np.array_equal(b, M.split_index(w)[0])
            d = zlib.decompressobj(-15)
            assert d.decompress(b.tobytes()) + d.flush() == c.tobytes() and d.unused_data == b"" and d.eof
        outs, ires, err = G.gpu_inflate_chunks(dev1, bare, [c.size for c in chunks])
        assert err is None and (ires["status"] == 0).all()
        for c, o in zip(chunks, outs):
            assert np.array_equal(o, c)
    finally:
        dev0.close()
        dev1.close()


def test_oversized_chunk_is_rejected_at_the_c_abi(cuda_device):
    """A deflate op longer than the largest segment (1 MiB) is refused before anything is launched."""
    dev = G.open_device(SEG)
    try:
        big = torch.zeros((1 << 20) + 4096, dtype=torch.uint8, device="cuda")
        dst = torch.zeros((1 << 20) + 65536, dtype=torch.uint8, device="cuda")
        ops = np.zeros(2, capi.CHUNK_DTYPE)
        ops["src"] = big.data_ptr()
        ops["dst"] = dst.data_ptr()
        ops["src_len"] = [1024, (1 << 20) + 1]
        ops["dst_cap"] = 600000
        with pytest.raises(capi.BitarError) as ei:
            dev.enqueue("deflate", 0, ops)
        assert ei.value.code == capi.E_INVALID
        ops["src_len"] = [1024, 1 << 20]          # the largest segment itself is fine
        ops["dst_cap"] = [4096, (1 << 20) + 4096]
        ops["dst"][1] = dst.data_ptr() + 8192
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        assert (res["status"] == 0).all()
    finally:
        dev.close()


def test_one_call_mixes_host_and_device_buffers(cuda_device):
    """Staging is decided per op: a Decompress() whose compressed buffers are partly pinned host memory and partly
    device memory, into destinations of both kinds, gives the right bytes everywhere."""
    L = capi.lib()
    data = synth.lineitem_like(8 * SEG)
    chunks = _chunks(data)
    dev = G.open_device(SEG)
    try:
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        n = len(chunks)
        stride = 70000
        import ctypes as C
        hp = C.c_void_p()
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, 2 * n * stride, 64, C.byref(hp)))
        host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), (2 * n * stride,))
        host[:] = 0xA5
        d_in = torch.zeros(n * stride, dtype=torch.uint8, device="cuda")
        d_out = torch.full((n * stride,), 0xA5, dtype=torch.uint8, device="cuda")
        ops = np.zeros(n, capi.CHUNK_DTYPE)
        for i, z in enumerate(comps):
            if i % 2:   # compressed bytes in pinned host memory
                host[i * stride:i * stride + z.size] = z
                ops["src"][i] = hp.value + i * stride
            else:
                d_in[i * stride:i * stride + z.size] = torch.from_numpy(z).cuda()
                ops["src"][i] = d_in.data_ptr() + i * stride
            ops["src_len"][i] = z.size
            # destinations: 0,1 device; 2,3 host; ...
            ops["dst"][i] = (hp.value + (n + i) * stride) if (i // 2) % 2 else (d_out.data_ptr() + i * stride)
            ops["dst_cap"][i] = chunks[i].size
        torch.cuda.synchronize()
        r = dev.enqueue("inflate", 0, ops)
        dev.wait(0)
        assert (r["status"] == 0).all() and [int(x) for x in r["produced"]] == [c.size for c in chunks]
        dh = d_out.cpu().numpy()
        for i, c in enumerate(chunks):
            got = host[(n + i) * stride:(n + i) * stride + c.size] if (i // 2) % 2 else dh[i * stride:i * stride + c.size]
            assert np.array_equal(got, c), i
        capi.check(L.bitar_mem_free(capi.MEM_PINNED, 0, hp))
    finally:
        dev.close()


@pytest.mark.parametrize("k,seg", [(4, 59460), (16, 59460), (3, 4096)])
def test_chained_segments(cuda_device, k, seg):
    """max_sgl_segs = k (/root/reference/src/include/config.h:90-96, src/memory.cc:394-398,420-424): k segments are ONE
    operation, i.e. one DEFLATE stream over k * S bytes whose output spans the k slots taken for them; Compress() lists a
    buffer per slot that holds data, Decompress() puts the streams together again; zlib inflates every stream; every slot
    comes back."""
    from bitar_b200.engine import Buf, sgl_join
    data = synth.lineitem_like(37 * seg + 1234)
    dev = G.open_device(seg, max_sgl_segs=k, max_preallocate_memzones=64)
    try:
        assert dev.sgl == k
        free0 = dev.slots_free()
        src = G.to_dev(data)
        bufs = dev.Compress(0, Buf(src.data_ptr(), data.size))
        streams = sgl_join(bufs, dev.slot)
        assert len(streams) == (data.size + k * seg - 1) // (k * seg)
        assert all(b.size <= dev.slot for b in bufs)
        for g, (ptr, n) in enumerate(streams):
            t = torch.empty(n, dtype=torch.uint8, device="cuda")
            capi.check(capi.lib().bitar_qp_memcpy(dev._h, 0, t.data_ptr(), ptr, n))
            dev.wait(0)
            part = data[g * k * seg:(g + 1) * k * seg]
            assert np.array_equal(O.inflate_chunk(t.cpu().numpy(), part.size), part), g
        out = torch.full((len(streams) * k * seg + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        total = dev.Decompress(0, bufs, Buf(out.data_ptr(), len(streams) * k * seg))
        assert total == data.size and np.array_equal(out[:total].cpu().numpy(), data)
        dev.Recycle(bufs)
        assert dev.slots_free() == free0
        # a second round finds contiguous groups again
        bufs = dev.Compress(0, Buf(src.data_ptr(), data.size))
        assert len(sgl_join(bufs, dev.slot)) == len(streams)
        dev.Recycle(bufs)
        assert dev.slots_free() == free0
    finally:
        dev.close()


def test_failed_op_of_a_staged_call_never_exposes_stale_stage_bytes(cuda_device):
    """Decompress() with everything in pinned host memory is staged through device memory and copied back in whole
    segments (one copy-engine transfer per batch).  When an op in the middle fails, the caller's segment may be left as
    it was or zeroed -- it must never receive bytes that an EARLIER call left in the stage."""
    import ctypes as C
    L = capi.lib()
    seg = 59460
    data = synth.lineitem_like(12 * seg)
    n = data.size // seg
    dev = G.open_device(seg, slot_mem_kind=capi.MEM_PINNED, max_preallocate_memzones=n + 8)
    h_in, h_out = C.c_void_p(), C.c_void_p()
    try:
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, data.size, 64, C.byref(h_in)))
        capi.check(L.bitar_mem_alloc(capi.MEM_PINNED, 0, n * seg + 64, 64, C.byref(h_out)))
        C.memmove(h_in.value, data.ctypes.data, data.size)
        ops, slots = dev.compress_ops(h_in.value, data.size)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        back = np.ctypeslib.as_array(C.cast(h_out.value, C.POINTER(C.c_uint8)), shape=(n * seg + 64,))
        dev.enqueue("inflate", 0, dev.decompress_ops(slots, res["produced"], h_out.value))   # fills the stage with plaintext
        dev.wait(0)
        assert np.array_equal(back[:data.size], data)
        C.memset(h_out.value, 0xA5, n * seg + 64)
        slot5 = np.ctypeslib.as_array(C.cast(int(slots[5]), C.POINTER(C.c_uint8)), shape=(int(res["produced"][5]),))
        slot5[0] |= 0x06                                       # block type 3: no such block
        ires = dev.enqueue("inflate", 0, dev.decompress_ops(slots, res["produced"], h_out.value))
        with pytest.raises(capi.BitarError):
            dev.wait(0)
        assert int(ires["status"][5]) != 0 and (np.delete(ires["status"], 5) == 0).all()
        seg5 = back[5 * seg:6 * seg]
        assert ((seg5 == 0) | (seg5 == 0xA5)).all(), "the failed op's segment received stale bytes"
        for i in (4, 6):
            assert np.array_equal(back[i * seg:(i + 1) * seg], data[i * seg:(i + 1) * seg])
        assert sum(dev.put_slot(s) for s in slots[::-1]) == n
    finally:
        for b in (h_in, h_out):
            if b.value:
                L.bitar_mem_free(capi.MEM_PINNED, 0, b)
        dev.close()
