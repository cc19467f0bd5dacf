"""ctypes view of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/bitar_oracle.c header).  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

HUFFMAN_FIXED = 1
HUFFMAN_DYNAMIC = 2

RFC_ERRORS = {
    -1: "TRUNCATED", -2: "BTYPE", -3: "STORED_LEN", -4: "OUTPUT_FULL", -5: "BAD_CODE",
    -6: "OVERSUBSCRIBED", -7: "INCOMPLETE", -8: "BAD_LENGTHS", -9: "BAD_SYMBOL",
    -10: "DIST_TOO_FAR", -11: "NO_EOB",
}


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("bitar_oracle.c", "rfc1951.c", "Makefile")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p = C.c_void_p
        L.oracle_compressed_seg_size.restype = C.c_uint32
        L.oracle_compressed_seg_size.argtypes = [C.c_uint32]
        L.oracle_max_seg_size.restype = C.c_uint32
        L.oracle_min_seg_size.restype = C.c_uint32
        L.oracle_stored_bound.restype = C.c_uint32
        L.oracle_stored_bound.argtypes = [C.c_uint32]
        L.oracle_deflate_chunk.restype = C.c_long
        L.oracle_deflate_chunk.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.oracle_inflate_chunk.restype = C.c_long
        L.oracle_inflate_chunk.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_int]
        L.oracle_crc32.restype = C.c_uint32
        L.oracle_crc32.argtypes = [u8p, C.c_size_t]
        L.oracle_adler32.restype = C.c_uint32
        L.oracle_adler32.argtypes = [u8p, C.c_size_t]
        L.oracle_compress_buffer.restype = C.c_int
        L.oracle_compress_buffer.argtypes = [u8p, C.c_uint64, C.c_uint32, u8p, C.c_uint32, u8p,
                                             C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_decompress_buffer.restype = C.c_int
        L.oracle_decompress_buffer.argtypes = [u8p, C.c_uint32, u8p, C.c_uint64, u8p, C.c_uint32,
                                               u8p, C.c_int, C.c_int]
        L.oracle_zlib_version.restype = C.c_char_p
        L.oracle_now.restype = C.c_double
        L.rfc1951_inflate.restype = C.c_long
        L.rfc1951_inflate.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.rfc_crc32.restype = C.c_uint32
        L.rfc_crc32.argtypes = [u8p, C.c_size_t]
        L.rfc_adler32.restype = C.c_uint32
        L.rfc_adler32.argtypes = [u8p, C.c_size_t]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if not isinstance(a, np.ndarray) else a)
    assert a.dtype == np.uint8
    return a


def _ptr(a):
    return a.ctypes.data if a.size else None


def compressed_seg_size(seg):
    return int(lib().oracle_compressed_seg_size(seg))


def stored_bound(n):
    return int(lib().oracle_stored_bound(n))


def deflate_chunk(data, level=1, window=15, huffman=HUFFMAN_DYNAMIC, cap=None):
    d = _u8(data)
    cap = cap if cap is not None else stored_bound(d.size) + 64
    out = np.empty(cap, np.uint8)
    r = lib().oracle_deflate_chunk(_ptr(d), d.size, _ptr(out), cap, level, window, huffman)
    if r < 0:
        raise RuntimeError(f"oracle deflate failed rc={r}")
    return out[:r].copy()


def inflate_chunk(comp, cap, window=15):
    c = _u8(comp)
    out = np.empty(max(cap, 1), np.uint8)
    r = lib().oracle_inflate_chunk(_ptr(c), c.size, _ptr(out), cap, window)
    if r < 0:
        raise RuntimeError(f"oracle inflate failed rc={r}")
    return out[:r].copy()


def rfc_inflate(comp, cap):
    """Independent decoder. Returns (bytes, info); raises ValueError with position on error."""
    c = _u8(comp)
    out = np.empty(max(cap, 1), np.uint8)
    consumed, blocks, err_bit = C.c_size_t(0), C.c_int(0), C.c_size_t(0)
    r = lib().rfc1951_inflate(_ptr(c), c.size, _ptr(out), cap, C.byref(consumed), C.byref(blocks),
                              C.byref(err_bit))
    if r < 0:
        raise ValueError(f"rfc1951 inflate: {RFC_ERRORS.get(r, r)} at bit {err_bit.value} "
                         f"(block {blocks.value})")
    return out[:r].copy(), {"consumed": consumed.value, "blocks": blocks.value}


def crc32(data):
    d = _u8(data)
    return int(lib().oracle_crc32(_ptr(d), d.size))


def adler32(data):
    d = _u8(data)
    return int(lib().oracle_adler32(_ptr(d), d.size))


def rfc_crc32(data):
    d = _u8(data)
    return int(lib().rfc_crc32(_ptr(d), d.size))


def rfc_adler32(data):
    d = _u8(data)
    return int(lib().rfc_adler32(_ptr(d), d.size))


def compress_buffer(data, seg, slot=None, level=1, window=15, huffman=HUFFMAN_DYNAMIC, threads=1):
    """bitar Compress() restated: returns (slots[n, slot] uint8, produced[n] uint32)."""
    d = _u8(data)
    n = (d.size + seg - 1) // seg
    slot = slot if slot is not None else max(compressed_seg_size(seg), stored_bound(seg))
    slots = np.zeros((n, slot), np.uint8)
    produced = np.zeros(n, np.uint32)
    rc = lib().oracle_compress_buffer(_ptr(d), d.size, seg, _ptr(slots), slot, _ptr(produced), level,
                                      window, huffman, threads)
    if rc:
        raise RuntimeError(f"oracle compress_buffer rc={rc}")
    return slots, produced


def decompress_buffer(slots, produced, seg, window=15, threads=1):
    """bitar Decompress() restated: segment i lands at i*seg; returns the trimmed output."""
    n, slot = slots.shape
    out = np.zeros(n * seg, np.uint8)
    got = np.zeros(n, np.uint32)
    rc = lib().oracle_decompress_buffer(_ptr(slots), slot, _ptr(produced), n, _ptr(out), seg,
                                        _ptr(got), window, threads)
    if rc:
        raise RuntimeError(f"oracle decompress_buffer rc={rc}")
    return out[:int(got.sum())], got
