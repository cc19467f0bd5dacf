"""GPU: the reference's object-level flow (apps/demo_app.cc:332-357, 487-548, 550-693) through the
host-side mirror: Compress -> Decompress -> memcmp -> Recycle, sync and async, multi queue pair."""
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
