"""GPU: the reference's object-level flow (apps/demo_app.cc:332-357, 487-548, 550-693) through the
host-side mirror: Compress -> Decompress -> memcmp -> Recycle, sync and async, multi queue pair."""
import zlib
import numpy as np
import pytest
import torch

import gpu_util as G
from bitar_b200 import _capi as capi
from bitar_b200 import engine as E
from bitar_b200 import synth

pytestmark = pytest.mark.gpu
SEG = 59460


def test_sync_flow_and_recycle(cuda_device):
    data = synth.lineitem_like(50 * SEG + 777)
    drv = E.CompressDriver.Instance()
    ids = drv.ListAvailableDeviceIds()
    dev = drv.GetDevices(ids[:1], 2)[0]
    assert dev.num_qps() == 2
    dev.Initialize(E.Configuration(decompressed_seg_size=SEG, max_preallocate_memzones=64))
    try:
        src = G.to_dev(data)
        free0 = dev.slots_free()
        bufs = dev.Compress(0, E.Buf(src.data_ptr(), data.size))
        assert len(bufs) == (data.size + SEG - 1) // SEG
        assert dev.slots_free() == free0 - len(bufs)
        out = torch.zeros(len(bufs) * SEG, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        total = dev.Decompress(1, bufs, E.Buf(out.data_ptr(), out.numel()))
        assert total == data.size and np.array_equal(out[:total].cpu().numpy(), data)
        assert dev.Recycle(bufs) == len(bufs)       # apps/demo_app.cc:288-290
        assert dev.Recycle(bufs) == 0               # second time: not occupied any more
        assert dev.slots_free() == free0
        # empty input is not an error (src/device.cc:161-164, 244-246)
        assert dev.Compress(0, E.Buf(src.data_ptr(), 0)) == [] and dev.Compress(0, None) == []
        assert dev.Decompress(0, [], None) == 0
        with pytest.raises(E.BitarError) as ei:      # CapacityError, src/device.cc:248-254
            dev.Decompress(0, [E.Buf(src.data_ptr(), 10)] * 3, E.Buf(out.data_ptr(), 2 * SEG))
        assert ei.value.code == capi.E_CAPACITY
        with pytest.raises(E.BitarError) as ei:      # Invalid qp, src/device.cc:451-454
            dev.Compress(5, E.Buf(src.data_ptr(), 100))
        assert ei.value.code == capi.E_INVALID
    finally:
        dev.close()


def test_pool_grows_on_demand(cuda_device):
    data = synth.lineitem_like(64 * 8192)
    dev = G.open_device(8192, max_preallocate_memzones=20)
    try:
        src = G.to_dev(data)
        bufs = dev.Compress(0, E.Buf(src.data_ptr(), data.size))   # needs 64 > 20 slots
        assert len(bufs) == 64 and dev.Recycle(bufs) == 64
    finally:
        dev.close()


def test_async_multi_queue_pair(cuda_device):
    """EvaluateAsync: input split evenly over queue pairs (apps/demo_app.cc:577-596); callbacks run off
    the caller's thread and return kAsyncReturnOK."""
    nqp = 4
    data = synth.lineitem_like(nqp * 30 * SEG)
    dev = E.CompressDevice(0, nqp).Initialize(E.Configuration(decompressed_seg_size=SEG, max_preallocate_memzones=256))
    try:
        src = G.to_dev(data)
        part = data.size // nqp
        results = {}

        def on_compressed(device_id, qp, bufs):
            results[qp] = bufs
            return E.kAsyncReturnOK if not isinstance(bufs, Exception) else 1
        calls = [E.CompressAsync(dev, qp, E.Buf(src.data_ptr() + qp * part, part), on_compressed) for qp in range(nqp)]
        assert [c.wait() for c in calls] == [E.kAsyncReturnOK] * nqp
        outs = [torch.zeros(len(results[qp]) * SEG, dtype=torch.uint8, device="cuda") for qp in range(nqp)]
        torch.cuda.synchronize()
        sizes = {}

        def on_decompressed(device_id, qp, status):
            sizes[qp] = status
            return E.kAsyncReturnOK if not isinstance(status, Exception) else 1
        calls = [E.DecompressAsync(dev, qp, results[qp], E.Buf(outs[qp].data_ptr(), outs[qp].numel()), on_decompressed)
                 for qp in range(nqp)]
        assert [c.wait() for c in calls] == [E.kAsyncReturnOK] * nqp
        for qp in range(nqp):
            assert sizes[qp] == part
            assert np.array_equal(outs[qp][:part].cpu().numpy(), data[qp * part:(qp + 1) * part])
            assert dev.Recycle(results[qp]) == len(results[qp])
    finally:
        dev.close()


def test_pinned_host_zero_copy(cuda_device):
    """Host buffers from the pinned pool are consumed and produced in place (the Rtememzone analogue)."""
    import ctypes as C
    data = synth.lineitem_like(20 * SEG)
    dev = G.open_device(SEG, slot_mem_kind=capi.MEM_PINNED, max_preallocate_memzones=32)
    try:
        p = C.c_void_p()
        capi.check(capi.lib().bitar_mem_alloc(capi.MEM_PINNED, 0, data.size, 64, C.byref(p)))
        C.memmove(p.value, data.ctypes.data, data.size)
        bufs = dev.Compress(0, E.Buf(p.value, data.size))
        q = C.c_void_p()
        capi.check(capi.lib().bitar_mem_alloc(capi.MEM_PINNED, 0, len(bufs) * SEG, 64, C.byref(q)))
        total = dev.Decompress(0, bufs, E.Buf(q.value, len(bufs) * SEG))
        back = np.ctypeslib.as_array(C.cast(q.value, C.POINTER(C.c_uint8)), shape=(total,))
        assert total == data.size and np.array_equal(back, data)
        # compressed bytes are host-readable: inflate one with zlib
        import zlib
        first = np.ctypeslib.as_array(C.cast(bufs[0].ptr, C.POINTER(C.c_uint8)), shape=(bufs[0].size,))
        assert zlib.decompressobj(-15).decompress(first.tobytes()) == data[:SEG].tobytes()
        dev.Recycle(bufs)
        capi.check(capi.lib().bitar_mem_free(capi.MEM_PINNED, 0, p))
        capi.check(capi.lib().bitar_mem_free(capi.MEM_PINNED, 0, q))
    finally:
        dev.close()


def test_chunks_frame_as_gzip_members(cuda_device):
    """the deflate kernel's chunks (index stripped) + its CRC-32s framed as gzip members are a
    multi-member gzip file that Python's gzip module (zlib) decompresses to the original buffer."""
    import gzip
    data = synth.lineitem_like(9 * SEG + 4321)
    dev = G.open_device(SEG, checksum_type=capi.CHECKSUM_CRC32)
    try:
        chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        blob = E.gzip_members(comps, res, [c.size for c in chunks])
        assert gzip.decompress(blob) == data.tobytes()
    finally:
        dev.close()


def test_chunks_frame_as_zlib_streams(cuda_device):
    """chunk + the kernel's Adler-32 framed per RFC 1950 is accepted (and verified) by zlib.decompress."""
    import zlib
    data = synth.lineitem_like(5 * SEG + 99)
    dev = G.open_device(SEG, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
        comps, res, err = G.gpu_deflate_chunks(dev, chunks)
        assert err is None
        for c, z in zip(chunks, E.zlib_streams(comps, res)):
            assert zlib.decompress(z) == c.tobytes()
    finally:
        dev.close()


def test_framed_foreign_streams_inflate_on_gpu(cuda_device):
    """zlib.compress() output (RFC 1950) and gzip members (RFC 1952) -- what Arrow's GZIP codec and
    Parquet pages hold -- inflate on the GPU once engine.unframe() has located the raw stream; the kernel's Adler-32 /
    CRC-32 equal the trailers."""
    import gzip
    data = synth.lineitem_like(7 * SEG + 1234)
    chunks = [data[o:o + SEG] for o in range(0, data.size, SEG)]
    framed = [np.frombuffer(zlib.compress(c.tobytes(), 6) if i % 2 else gzip.compress(c.tobytes(), 1), np.uint8)
              for i, c in enumerate(chunks)]
    raw, expect = [], []
    for f in framed:
        off, ln, kind, exp = E.unframe(f)
        raw.append(f[off:off + ln])
        expect.append((kind, exp))
    dev = G.open_device(SEG, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        outs, res, err = G.gpu_inflate_chunks(dev, raw, [c.size for c in chunks])
        assert err is None and (res["status"] == 0).all()
        for c, o, r, (kind, exp) in zip(chunks, outs, res, expect):
            assert np.array_equal(o, c)
            if kind == "zlib":
                assert int(r["checksum"]) >> 32 == exp[0]
            else:
                assert int(r["checksum"]) & 0xFFFFFFFF == exp[0] and c.size == exp[1]
    finally:
        dev.close()


def _crc32_combine(crc1, crc2, len2):
    """zlib's crc32_combine(): CRC of A||B from CRC(A), CRC(B), len(B) (GF(2) matrix method)."""
    def times(mat, vec):
        s, i = 0, 0
        while vec:
            if vec & 1:
                s ^= mat[i]
            vec >>= 1
            i += 1
        return s

    def square(mat):
        return [times(mat, mat[n]) for n in range(32)]

    if len2 == 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]
    even = square(odd)
    odd = square(even)
    while True:
        even = square(odd)
        if len2 & 1:
            crc1 = times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = square(even)
        if len2 & 1:
            crc1 = times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def _adler32_combine(a1, a2, len2):
    base = 65521
    rem = len2 % base
    s1 = a1 & 0xFFFF
    s2 = (rem * s1) % base
    s1 += (a2 & 0xFFFF) + base - 1
    s2 += ((a1 >> 16) & 0xFFFF) + ((a2 >> 16) & 0xFFFF) + base - rem
    if s1 >= base:
        s1 -= base
    if s1 >= base:
        s1 -= base
    if s2 >= (base << 1):
        s2 -= (base << 1)
    if s2 >= base:
        s2 -= base
    return s1 | (s2 << 16)


def test_full_size_round_trip_and_checksum_of_checksums(cuda_device):
    """BASELINE config 2 at full size (1 GiB, 18 059 segments of 59 460 bytes), checked through size-independent
    properties: Compress -> Decompress is the identity; the per-segment CRC-32 / Adler-32 the deflate kernel
    returns, combined in segment order, equal zlib's checksum of the whole buffer; the inflate kernel returns
    the same per-segment checksums; total compressed size within the tolerance of zlib level 1 (sample)."""
    import zlib
    import oracle_lib as O
    U = 1 << 30
    data = synth.lineitem_like(U)
    n = (U + SEG - 1) // SEG
    dev = G.open_device(SEG, max_preallocate_memzones=n + 8, checksum_type=capi.CHECKSUM_CRC32_ADLER32)
    try:
        src = torch.from_numpy(data).cuda()
        out = torch.empty(n * SEG + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ops, slots = dev.compress_ops(src.data_ptr(), U)
        res = dev.enqueue("deflate", 0, ops)
        dev.wait(0)
        assert (res["status"] == 0).all()
        crc, adler = 0, 1
        for c, ln in zip(res["checksum"], ops["src_len"]):
            crc = _crc32_combine(crc, int(c) & 0xFFFFFFFF, int(ln))
            adler = _adler32_combine(adler, int(c) >> 32, int(ln))
        assert crc == zlib.crc32(data) and adler == zlib.adler32(data)
        ires = dev.enqueue("inflate", 0, dev.decompress_ops(slots, res["produced"], out.data_ptr()))
        dev.wait(0)
        assert int(ires["produced"].sum()) == U and bool(torch.equal(out[:U], src))
        assert np.array_equal(ires["checksum"], res["checksum"])
        third, part = U // 3, (8 << 20) // SEG * SEG
        zs = np.concatenate([data[i * third:i * third + part] for i in range(3)])
        _, zp = O.compress_buffer(zs, SEG, threads=8)
        gpu_ratio, zlib_ratio = U / float(res["produced"].sum()), zs.size / float(zp.sum())
        assert gpu_ratio >= zlib_ratio / 1.05, (gpu_ratio, zlib_ratio)
        assert sum(dev.put_slot(s) for s in slots[::-1]) == n
    finally:
        dev.close()
