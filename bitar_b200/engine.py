"""Host-side mirror of bitar's operator interface over the C-ABI (tests + bench harness).

Names, argument meaning and error behaviour follow the reference's C++ API so the parity tests read
like the reference's own demo_app flow (/root/reference/apps/demo_app.cc:332-357, 487-548):

    CompressDriver.ListAvailableDeviceIds / GetDevices   src/include/driver.h:47-57
    CompressDevice.Initialize / Compress / Decompress /
                   Recycle / device_id / num_qps          src/include/device.h:83-134
    Configuration (BlueFieldConfiguration fields)        src/include/config.h:62-183
    CompressAsync / DecompressAsync                      src/include/util.h:216-236

The production host side is the C++ facade in bitar_b200/host; this module exists so that Python
tests and bench.py can drive the same library.  Buffers are device-accessible address ranges
(CUDA device memory or pinned host memory), described by ``Buf(ptr, size)``.
"""
import ctypes as C
import threading
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from ._capi import (CHECKSUM_ADLER32, CHECKSUM_CRC32, CHECKSUM_CRC32_ADLER32, CHECKSUM_NONE,  # noqa: F401
                    HUFFMAN_DYNAMIC, HUFFMAN_FIXED, MEM_DEVICE, MEM_PINNED, BitarError)

kAsyncReturnOK = 2  # src/include/util.h:45
kMinSegSize, kRefMaxSegSize, kMaxSegSize = 8, 59460, 1 << 20


@dataclass
class Buf:
    """A non-owning view (arrow::Buffer analogue): address + size in bytes."""
    ptr: int
    size: int


@dataclass
class Configuration:
    """Configuration<Class_CUDA> + BlueFieldConfiguration knobs (src/include/config.h:146-152,183)."""
    burst_size: int = 32
    max_sgl_segs: int = 1
    decompressed_seg_size: int = 2048        # kDefaultSegSize
    window_size: int = 0                     # 0 -> device max (15)
    huffman_enc: int = HUFFMAN_DYNAMIC
    max_preallocate_memzones: int = 2560     # RTE_MAX_MEMZONE
    checksum_type: int = CHECKSUM_NONE
    slot_mem_kind: int = MEM_DEVICE
    compressed_seg_size: int = 0             # derived unless set
    emit_index: bool = True                  # False: chunks are bare RFC 1951 streams (no parallel-inflate index)

    def resolved_compressed_seg_size(self):
        return self.compressed_seg_size or int(capi.lib().bitar_compressed_seg_size(self.decompressed_seg_size))


class CompressDevice:
    """One CUDA device with ``num_qps`` queue pairs (CUDA streams)."""

    def __init__(self, device_id, num_qps):
        self._device_id = int(device_id)
        self._num_qps = int(num_qps)
        self._h = None
        self.cfg = None
        self._keep = {}

    # -- lifecycle ------------------------------------------------------------------------------
    def Initialize(self, configuration: Configuration):
        c = capi.Cfg(configuration.decompressed_seg_size, configuration.compressed_seg_size,
                     configuration.max_preallocate_memzones, configuration.burst_size,
                     configuration.max_sgl_segs, configuration.window_size, configuration.huffman_enc,
                     configuration.checksum_type, configuration.slot_mem_kind, 0 if configuration.emit_index else 1)
        h = C.c_void_p()
        capi.check(capi.lib().bitar_dev_open(self._device_id, self._num_qps, C.byref(c), C.byref(h)))
        self._h = h
        out = capi.Cfg()
        capi.check(capi.lib().bitar_dev_config(h, C.byref(out)))
        self.cfg = out
        return self

    def close(self):
        if self._h is not None:
            capi.lib().bitar_dev_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_id(self):
        return self._device_id

    def num_qps(self):
        return self._num_qps

    @property
    def seg(self):
        return int(self.cfg.decompressed_seg_size)

    @property
    def slot(self):
        return int(self.cfg.compressed_seg_size)

    @property
    def sgl(self):
        """Segments chained into one stream (max_sgl_segs, src/include/config.h:90-96)."""
        return max(1, int(self.cfg.max_sgl_segs))

    def _guard(self):
        if self._h is None:
            raise BitarError(capi.E_INVALID, f"Compress device {self._device_id} has not started")

    # -- raw op interface (arrays of bitar_chunk / bitar_result) -------------------------------------
    def enqueue(self, kind, qp, ops, results=None):
        """Enqueue ops (numpy CHUNK_DTYPE array) on queue pair qp; returns the results array, which is
        filled when wait(qp) returns."""
        self._guard()
        ops = np.ascontiguousarray(ops, dtype=capi.CHUNK_DTYPE)
        if results is None:
            results = np.zeros(ops.size, capi.RESULT_DTYPE)
            results["status"] = 0xFFFFFFFF
        fn = capi.lib().bitar_qp_deflate if kind == "deflate" else capi.lib().bitar_qp_inflate
        capi.check(fn(self._h, qp, ops.ctypes.data if ops.size else None, ops.size,
                      results.ctypes.data if results.size else None))
        self._keep[qp] = (ops, results)
        return results

    def wait(self, qp):
        self._guard()
        capi.check(capi.lib().bitar_qp_wait(self._h, qp))

    def busy(self, qp):
        return bool(capi.lib().bitar_qp_busy(self._h, qp))

    def last_ms(self, qp):
        k, t = C.c_float(), C.c_float()
        capi.check(capi.lib().bitar_qp_last_ms(self._h, qp, C.byref(k), C.byref(t)))
        return k.value, t.value

    def stream(self, qp):
        return capi.lib().bitar_qp_stream(self._h, qp)

    def take_slots(self, n):
        arr = np.zeros(n, np.uint64)
        capi.check(capi.lib().bitar_slot_take_n(self._h, n, arr.ctypes.data if n else None))
        return arr

    def put_slot(self, ptr):
        return int(capi.lib().bitar_slot_put(self._h, C.c_void_p(int(ptr))))

    def put_slots(self, ptrs):
        """Recycle an array of slot addresses in one call; returns the number recycled."""
        arr = np.ascontiguousarray(ptrs, dtype=np.uint64)
        return int(capi.lib().bitar_slot_put_n(self._h, arr.ctypes.data if arr.size else None, arr.size))

    def slots_free(self):
        return int(capi.lib().bitar_slots_free(self._h))

    # -- array-level Compress/Decompress (what bench.py times) -------------------------------------------
    def compress_ops(self, src_ptr, nbytes, slots=None):
        """Op list of Compress(): segment i = [i*S, min((i+1)*S, size)) -> slot i (src/memory.cc:350-430)."""
        S, k = self.seg, self.sgl
        n = (nbytes + S - 1) // S
        if slots is None:
            slots = self.take_slots(n)
        if k > 1:
            # chained segments (src/memory.cc:350-430 with max_sgl_segs > 1): one op = one stream over k segments, its
            # destination the k slots taken for them, which must be one contiguous range
            g = (n + k - 1) // k
            cnt = np.minimum(k, n - np.arange(g) * k)
            first = slots[::k]
            expect = np.repeat(first, k)[:n] + (np.arange(n, dtype=np.uint64) % np.uint64(k)) * np.uint64(self.slot)
            if not np.array_equal(expect, slots):
                self.put_slots(slots)
                raise BitarError(capi.E_IO_ERROR, "the output slots of a chained operation are not contiguous")
            ops = np.zeros(g, capi.CHUNK_DTYPE)
            off = np.arange(g, dtype=np.uint64) * np.uint64(k * S)
            ops["src"] = np.uint64(src_ptr) + off
            ops["src_len"] = np.minimum(np.uint64(k * S), np.uint64(nbytes) - off).astype(np.uint32)
            ops["dst"] = first
            ops["dst_cap"] = (cnt * self.slot).astype(np.uint32)
            return ops, slots
        ops = np.zeros(n, capi.CHUNK_DTYPE)
        off = np.arange(n, dtype=np.uint64) * np.uint64(S)
        ops["src"] = np.uint64(src_ptr) + off
        ops["src_len"] = np.minimum(np.uint64(S), np.uint64(nbytes) - off).astype(np.uint32)
        ops["dst"] = slots
        ops["dst_cap"] = self.slot
        return ops, slots

    def decompress_ops(self, comp_ptrs, comp_lens, out_ptr):
        """Op list of Decompress(): op i inflates buffers[i] to out + i*S, capacity S (src/memory.cc:432-505)."""
        n = len(comp_ptrs)
        ops = np.zeros(n, capi.CHUNK_DTYPE)
        ops["src"] = comp_ptrs
        ops["src_len"] = comp_lens
        ops["dst"] = np.uint64(out_ptr) + np.arange(n, dtype=np.uint64) * np.uint64(self.seg)
        ops["dst_cap"] = self.seg
        return ops

    # -- the reference's object-level API -------------------------------------------------------------------
    def Compress(self, queue_pair_id, decompressed_buffer):
        """-> list[Buf] of compressed segments in input order; views into pool slots that the caller must
        Recycle() (src/device.cc:156-238).  None/empty input -> [] (src/device.cc:161-164)."""
        if decompressed_buffer is None or decompressed_buffer.size == 0:
            return []
        self._guard()
        self._entry_guard(queue_pair_id)
        ops, slots = self.compress_ops(decompressed_buffer.ptr, decompressed_buffer.size)
        try:
            res = self.enqueue("deflate", queue_pair_id, ops)
            self.wait(queue_pair_id)
        except BitarError:
            for s in slots[::-1]:   # ReleaseAll, src/device.cc:537-542
                self.put_slot(s)
            raise
        self.last_results = res
        if self.sgl > 1:
            bufs, unused = sgl_split(slots, res["produced"], self.sgl, self.slot)
            self.put_slots(unused)       # the slots a stream did not reach (the reference leaves them occupied)
            return bufs
        return [Buf(int(p), int(n)) for p, n in zip(slots, res["produced"])]

    def Decompress(self, queue_pair_id, compressed_buffers, decompressed_buffer):
        """Inflate buffers[i] to decompressed_buffer.ptr + i*S; returns the total size (the reference
        Resize()s the ResizableBuffer, src/device.cc:315).  decompressed_buffer.size is its capacity."""
        if not compressed_buffers:
            return 0
        if self.sgl > 1:
            streams = sgl_join(compressed_buffers, self.slot)
            need = len(streams) * self.sgl * self.seg
        else:
            need = len(compressed_buffers) * self.seg
        if decompressed_buffer is None or decompressed_buffer.size < need:
            raise BitarError(capi.E_CAPACITY, f"The decompressed_buffer is required to be >= {need} bytes")
        self._guard()
        self._entry_guard(queue_pair_id)
        if self.sgl > 1:
            ops = np.zeros(len(streams), capi.CHUNK_DTYPE)
            ops["src"] = [p for p, _ in streams]
            ops["src_len"] = [n for _, n in streams]
            ops["dst"] = np.uint64(decompressed_buffer.ptr) + np.arange(len(streams), dtype=np.uint64) * np.uint64(self.sgl * self.seg)
            ops["dst_cap"] = self.sgl * self.seg
        else:
            ops = self.decompress_ops(np.array([b.ptr for b in compressed_buffers], np.uint64),
                                      np.array([b.size for b in compressed_buffers], np.uint32),
                                      decompressed_buffer.ptr)
        res = self.enqueue("inflate", queue_pair_id, ops)
        self.wait(queue_pair_id)
        self.last_results = res
        return int(res["produced"].sum())

    def Recycle(self, buffers):
        """Returns the number of buffers recycled, walking in reverse (src/device.cc:320-327)."""
        return self.put_slots(np.array([b.ptr for b in buffers], np.uint64))

    def _entry_guard(self, qp):  # src/device.cc:443-462
        if qp >= self._num_qps:
            raise BitarError(capi.E_INVALID, f"queue_pair_id must be in the range of [0, {self._num_qps})")
        if self.busy(qp):
            raise BitarError(capi.E_CANCELLED, f"Queue pair {qp} of compress device {self._device_id} is busy")


def sgl_split(slots, produced, k, slot):
    """Chained segments, compress side: the buffers of every stream as the reference's dequeue callback lists them
    (src/device.cc:183-195) -- one per destination slot that holds data, all of them full but the last.  A stream that
    ends exactly at a slot boundary gets an empty buffer after it, so that the end of a stream is always a buffer shorter
    than a slot (what sgl_join() goes by).  Returns (buffers, slots the streams did not reach)."""
    bufs, unused = [], []
    n = len(slots)
    for g, p in enumerate(produced):
        p = int(p)
        cnt = min(k, n - g * k)
        used = (p + slot - 1) // slot
        for j in range(used):
            bufs.append(Buf(int(slots[g * k + j]), min(slot, p - j * slot)))
        if p % slot == 0:
            bufs.append(Buf(int(slots[g * k]) + p, 0))
            if used < cnt:
                used += 1     # the empty buffer sits on (and keeps) the next slot
        unused.extend(int(x) for x in slots[g * k + used:g * k + cnt])
    return bufs, np.array(unused, np.uint64)


def sgl_join(buffers, slot):
    """Chained segments, decompress side (src/memory.cc:432-505 with max_sgl_segs > 1): the buffers of one stream are
    consecutive full slots and end with a shorter one; they must be one contiguous range.  Returns [(address, bytes)]
    per stream."""
    streams, i = [], 0
    while i < len(buffers):
        first, total = buffers[i].ptr, 0
        while True:
            b = buffers[i]
            if b.ptr != first + total:
                raise BitarError(capi.E_INVALID, "the compressed buffers of a chained operation are not contiguous")
            total += b.size
            i += 1
            if b.size < slot:
                break
            if i == len(buffers):
                raise BitarError(capi.E_INVALID, "the last chained operation has no end (a buffer shorter than a slot)")
        streams.append((first, total))
    return streams


class CompressDriver:
    """CompressDriver<Class_CUDA> (src/include/driver.h:40-66)."""
    _instance = None

    @classmethod
    def Instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def ListAvailableDeviceIds(self):
        n = capi.lib().bitar_cuda_device_count()
        if n == 0:
            raise BitarError(capi.E_INVALID, "No compress device is available with driver name: CUDA")
        return list(range(n))

    def GetDevices(self, device_ids, num_workers=None):
        """Spread ``num_workers`` queue pairs over the devices as evenly as possible, each device at
        least one; the first W % D devices get one more (src/driver.cc:100-158, 192-223)."""
        avail = self.ListAvailableDeviceIds()
        for d in device_ids:
            if d not in avail:
                raise BitarError(capi.E_INVALID, f"Device id {d} is not available")
        if num_workers is None:
            num_workers = len(device_ids)
        if num_workers == 0:
            raise BitarError(capi.E_INVALID, "Not enough worker lcores for setting up queue pairs for devices.")
        if len(device_ids) > num_workers:
            raise BitarError(capi.E_INVALID, f"The number of devices to set up ({len(device_ids)}) is greater than "
                                            f"the number of available worker lcores ({num_workers}).")
        base, rem = divmod(num_workers, len(device_ids))
        return [CompressDevice(d, base + (1 if i < rem else 0)) for i, d in enumerate(device_ids)]


# -- framing: the chunks as members of a gzip file --------------------------------------
INDEX_MAGIC = 0xB17A0B02


def stream_length(chunk):
    """Bytes of the raw DEFLATE stream at the start of a compressed chunk (numpy uint8 array): the chunk minus the
    parallel-inflate index, when it carries one (bitar_b200/csrc/deflate_common.h)."""
    c = np.ascontiguousarray(chunk, dtype=np.uint8)
    if c.size >= 16 and int(c[-4:].view("<u4")[0]) == INDEX_MAGIC:
        total, end_bit = int(c[-8:-4].view("<u4")[0]), int(c[-12:-8].view("<u4")[0])
        full, rem = divmod(total, 65536)
        entries = full * 33 + ((1 + (rem + 2047) // 2048) if rem else 0)
        if (end_bit + 7) // 8 + 4 * (entries + 3) == c.size:
            return (end_bit + 7) // 8
    return c.size


def strip_index(chunk):
    """The raw DEFLATE stream of a compressed chunk (numpy uint8 array), without the parallel-inflate index."""
    c = np.ascontiguousarray(chunk, dtype=np.uint8)
    return c[:stream_length(c)]


def gzip_members(chunks, results, sizes):
    """RFC 1952 framing: every compressed chunk becomes one gzip member (header, raw DEFLATE stream, CRC-32,
    ISIZE); the concatenation is a valid multi-member gzip file that gzip / zlib tools decompress to the original
    buffer.  `results` are the Compress() results of a device configured with a CRC-32 checksum, `sizes` the
    uncompressed segment sizes."""
    import struct
    out = bytearray()
    for c, r, n in zip(chunks, results, sizes):
        c = np.ascontiguousarray(c, dtype=np.uint8)
        out += b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff"
        out += c[:stream_length(c)].tobytes()
        out += struct.pack("<II", int(r["checksum"]) & 0xFFFFFFFF, int(n) & 0xFFFFFFFF)
    return bytes(out)


def zlib_streams(chunks, results):
    """RFC 1950 framing: every compressed chunk becomes one zlib stream (CMF/FLG 78 01: 32 KiB window, fastest
    level, no dictionary; raw DEFLATE stream; Adler-32 big-endian) -- what Arrow's / Parquet's ZLIB-format readers
    and `zlib.decompress` expect per buffer.  `results` are the Compress() results of a device configured with an
    Adler-32 checksum.  Returns a list of bytes objects, one per chunk."""
    import struct
    out = []
    for c, r in zip(chunks, results):
        c = np.ascontiguousarray(c, dtype=np.uint8)
        out.append(b"\x78\x01" + c[:stream_length(c)].tobytes() + struct.pack(">I", int(r["checksum"]) >> 32))
    return out


def unframe(buf):
    """The other direction of zlib_streams() / gzip_members(): locate the raw DEFLATE stream inside ONE RFC 1950
    (zlib) stream or ONE RFC 1952 (gzip) member, as Arrow's / Parquet's GZIP pages and `zlib.compress` produce them,
    so that Decompress() can be pointed at it.  Returns (offset, length, kind, expected): `kind` is "zlib" or "gzip",
    `expected` the trailer's checksum (Adler-32 for zlib, CRC-32 for gzip; plus ISIZE for gzip as a second element) to
    compare with the checksum the inflate kernel returns.  Raises ValueError for anything else (preset dictionaries,
    unknown methods, truncated framing).  The DEFLATE stream's own end is found by the decoder, not here: `length`
    spans up to the trailer."""
    import struct
    b = np.ascontiguousarray(buf, dtype=np.uint8)
    n = b.size
    if n >= 18 and b[0] == 0x1F and b[1] == 0x8B:
        if b[2] != 8:
            raise ValueError("gzip member: method is not DEFLATE")
        flg, at = int(b[3]), 10
        if flg & 0xE0:
            raise ValueError("gzip member: reserved flag bits set")
        if flg & 4:                                   # FEXTRA
            if at + 2 > n:
                raise ValueError("gzip member: truncated header")
            at += 2 + int(b[at]) + 256 * int(b[at + 1])
        for bit in (8, 16):                           # FNAME, FCOMMENT: zero-terminated
            if flg & bit:
                while at < n and b[at] != 0:
                    at += 1
                at += 1
        if flg & 2:                                   # FHCRC
            at += 2
        if at + 8 > n:
            raise ValueError("gzip member: truncated")
        crc, isize = struct.unpack("<II", b[n - 8:].tobytes())
        return at, n - 8 - at, "gzip", (crc, isize)
    if n >= 6 and (int(b[0]) & 0x0F) == 8 and ((int(b[0]) << 8) | int(b[1])) % 31 == 0:
        if (int(b[0]) >> 4) > 7:
            raise ValueError("zlib stream: window larger than 32 KiB")
        if int(b[1]) & 0x20:
            raise ValueError("zlib stream: preset dictionary")
        (adler,) = struct.unpack(">I", b[n - 4:].tobytes())
        return 2, n - 6, "zlib", (adler,)
    raise ValueError("neither a zlib stream nor a gzip member")


def shard_range(n_chunks, rank, world):
    """Chunk range [first, last) of rank `rank` out of `world`: contiguous ranges of
    ceil(n / world) chunks, so that concatenating the ranks' outputs in rank order is the output of one
    Compress() over the whole buffer (apps/demo_app.cc:577-607 splits the same way over (device, qp))."""
    per = (n_chunks + world - 1) // world
    first = min(n_chunks, rank * per)
    return first, min(n_chunks, first + per)


def distribute_workers(num_workers, num_devices):
    """The queue-pair distribution rule alone (host logic, testable without a GPU)."""
    base, rem = divmod(num_workers, num_devices)
    return [base + (1 if i < rem else 0) for i in range(num_devices)]


# -- async (src/include/util.h:47-101, 216-236) ------------------------------------------------------------
_CB = C.CFUNCTYPE(None, C.c_void_p)


class _AsyncCall:
    """Completion through the queue pair's stream: the callback runs on a CUDA driver thread, the
    analogue of the worker lcore (src/include/util.h:133-151); its int result is read with wait()."""

    def __init__(self, device, qp, finish):
        self.device, self.qp = device, qp
        self.ret = 0          # 0 == never ran (apps/demo_app.cc:267-274)
        self.done = threading.Event()

        def _run(_arg):
            try:
                self.ret = finish()
            except Exception:  # pragma: no cover - surfaced through ret
                self.ret = 1   # EXIT_FAILURE
            self.done.set()
        self._cb = _CB(_run)

    def launch(self):
        capi.check(capi.lib().bitar_qp_on_complete(self.device._h, self.qp, C.cast(self._cb, C.c_void_p), None))

    def wait(self):
        """rte_eal_wait_lcore analogue: returns the callback's int."""
        self.done.wait()
        return self.ret


def CompressAsync(device, queue_pair_id, decompressed_buffer, result_callback):
    """result_callback(device_id, queue_pair_id, buffers_or_exception) -> int.  Returns the launched
    call (0 from the reference) or raises Cancelled when the queue pair is busy (-EBUSY there)."""
    device._entry_guard(queue_pair_id)
    if decompressed_buffer is None or decompressed_buffer.size == 0:
        call = _AsyncCall(device, queue_pair_id, lambda: result_callback(device.device_id(), queue_pair_id, []))
        call.launch()
        return call
    ops, slots = device.compress_ops(decompressed_buffer.ptr, decompressed_buffer.size)
    res = device.enqueue("deflate", queue_pair_id, ops)

    def finish():
        if (res["status"] != 0).any():
            for s in slots[::-1]:
                device.put_slot(s)
            return result_callback(device.device_id(), queue_pair_id,
                                   BitarError(capi.E_IO_ERROR, "compression operation failed"))
        return result_callback(device.device_id(), queue_pair_id,
                               [Buf(int(p), int(n)) for p, n in zip(slots, res["produced"])])
    call = _AsyncCall(device, queue_pair_id, finish)
    call.launch()
    return call


def DecompressAsync(device, queue_pair_id, compressed_buffers, decompressed_buffer, result_callback):
    """result_callback(device_id, queue_pair_id, status) -> int, status = total size or a BitarError."""
    device._entry_guard(queue_pair_id)
    need = len(compressed_buffers) * device.seg
    if compressed_buffers and (decompressed_buffer is None or decompressed_buffer.size < need):
        raise BitarError(capi.E_CAPACITY, f"The decompressed_buffer is required to be >= {need} bytes")
    ops = device.decompress_ops(np.array([b.ptr for b in compressed_buffers], np.uint64),
                                np.array([b.size for b in compressed_buffers], np.uint32),
                                decompressed_buffer.ptr if decompressed_buffer else 0)
    res = device.enqueue("inflate", queue_pair_id, ops)

    def finish():
        if (res["status"] != 0).any():
            return result_callback(device.device_id(), queue_pair_id,
                                   BitarError(capi.E_IO_ERROR, "decompression operation failed"))
        return result_callback(device.device_id(), queue_pair_id, int(res["produced"].sum()))
    call = _AsyncCall(device, queue_pair_id, finish)
    call.launch()
    return call
