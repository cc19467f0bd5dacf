"""ctypes view of tools/model/libmodel.so -- the kernels' host/device-shared code built for the CPU.

TEST TOOL: lets the exact decoder source of the inflate kernel (G = 1) and the sequential model of
the deflate kernel run without a GPU.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "model")
_LIB = os.path.join(_DIR, "libmodel.so")
STATUS = {0: "OK", 1: "OUT_OF_SPACE", 2: "DATA_ERROR", 3: "TRUNCATED"}
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s", "libmodel.so"], stderr=subprocess.DEVNULL)
        L = C.CDLL(_LIB)
        L.host_inflate_chunk.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.host_inflate_lane.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.model_deflate_chunk.restype = C.c_long
        L.model_deflate_chunk.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        L.model_selfcheck.restype = C.c_int
        _lib = L
    return _lib


def host_inflate(comp, cap, lbits=10, misalign=0, lane=False):
    """Run inflate_core.h (G=1) or, with lane=True, the lane-per-chunk state machine of inflate_lane.h.
    Returns (bytes, dict(status, produced, consumed, blocks))."""
    c = np.ascontiguousarray(comp, dtype=np.uint8)
    buf = np.full(cap + 64, 0xA5, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16 + misalign
    res = np.zeros(4, np.uint32)
    fn = lib().host_inflate_lane if lane else lib().host_inflate_chunk
    fn(c.ctypes.data if c.size else None, c.size, base + off, cap, res.ctypes.data, lbits)
    out = buf[off:off + int(res[0])].copy()
    guard_ok = bool((buf[:off] == 0xA5).all() and (buf[off + cap:] == 0xA5).all())
    return out, {"produced": int(res[0]), "status": int(res[1]), "consumed": int(res[2]),
                 "blocks": int(res[3]), "guard_ok": guard_ok}


def model_deflate(data, huffman=2, block=0):
    d = np.ascontiguousarray(data, dtype=np.uint8)
    cap = d.size + d.size // 8 + 1024
    out = np.empty(cap, np.uint8)
    r = lib().model_deflate_chunk(d.ctypes.data if d.size else None, d.size, out.ctypes.data, cap, huffman, block)
    if r < 0:
        raise RuntimeError("model deflate overflow")
    return out[:r].copy()
