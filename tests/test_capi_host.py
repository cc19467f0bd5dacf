"""CPU: the C-ABI library loads, exports every symbol include/bitar_cuda.h declares, and its pure
host logic (segment-size rule, error conventions without a device) behaves like the reference.
No compute calls: there is no GPU here."""
import ctypes as C
import os
import re

import pytest

import oracle_lib as O
from bitar_b200 import _capi as capi
from bitar_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bitar_cuda.h")).read()
    declared = sorted(set(re.findall(r"BITAR_API[^;(]*?\b(bitar_\w+)\s*\(", hdr)))
    assert declared == sorted(capi.EXPORTS)
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    assert C.sizeof(capi.Cfg) == 24 and C.sizeof(capi.DevInfo) == 104
    assert capi.CHUNK_DTYPE.itemsize == 24 and capi.RESULT_DTYPE.itemsize == 16


def test_compressed_seg_size_rule():
    L = capi.lib()
    for seg in (8, 100, 2048, 4095, 4096, 32768, 59460):
        assert L.bitar_reference_compressed_seg_size(seg) == O.compressed_seg_size(seg)   # src/config.cc:59-73
        assert L.bitar_compressed_seg_size(seg) >= O.stored_bound(seg)
    assert L.bitar_compressed_seg_size(59460) == 65406
    assert L.bitar_compressed_seg_size(4095) == 4100          # reference value 4096 cannot hold a stored block
    assert L.bitar_compressed_seg_size(65536) == max(int(65536 * 1.1), 65546)
    assert L.bitar_compressed_seg_size(1 << 20) >= (1 << 20) + 5 * 17


def test_no_device_errors_are_loud():
    L = capi.lib()
    if L.bitar_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.bitar_dev_open(0, 1, None, C.byref(h))
    assert rc == capi.E_INVALID and L.bitar_last_error()
    with pytest.raises(E.BitarError):
        E.CompressDriver.Instance().ListAvailableDeviceIds()   # "No compress device is available", src/driver.cc:184-187


def test_worker_distribution_rule():
    """src/driver.cc:103-117,152-155: min = W / D, the first W % D devices get one more."""
    assert E.distribute_workers(7, 2) == [4, 3]
    assert E.distribute_workers(8, 8) == [1] * 8
    assert E.distribute_workers(19, 8) == [3, 3, 3, 2, 2, 2, 2, 2]
    assert sum(E.distribute_workers(31, 8)) == 31


def test_chained_segment_buffer_lists():
    """The host logic of max_sgl_segs > 1 (src/memory.cc:394-398,471-475): a stream's buffers are full slots and a
    shorter last one, an empty buffer marks a stream that ends at a slot boundary, unused slots are reported back."""
    import numpy as np
    from bitar_b200.engine import BitarError, Buf, sgl_join, sgl_split
    slot, k = 1000, 4
    slots = np.array([10000 + i * slot for i in range(10)], np.uint64)
    bufs, unused = sgl_split(slots, [2500, 1000, 37], k, slot)
    assert [(b.ptr, b.size) for b in bufs] == [(10000, 1000), (11000, 1000), (12000, 500), (14000, 1000), (15000, 0), (18000, 37)]
    assert list(unused) == [13000, 16000, 17000, 19000]
    assert sgl_join(bufs, slot) == [(10000, 2500), (14000, 1000), (18000, 37)]
    bufs, unused = sgl_split(slots[:4], [4000], k, slot)          # every slot full: the end marker has no slot of its own
    assert [(b.ptr, b.size) for b in bufs][-1] == (14000, 0) and len(unused) == 0 and sgl_join(bufs, slot) == [(10000, 4000)]
    import pytest
    with pytest.raises(BitarError):
        sgl_join([Buf(10000, 1000), Buf(12000, 10)], slot)          # a gap
    with pytest.raises(BitarError):
        sgl_join([Buf(10000, 1000)], slot)                          # no end
