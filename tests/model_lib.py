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
        L.host_inflate_fast.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.host_inflate_indexed.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.host_inflate_spec.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.model_deflate_chunk.restype = C.c_long
        L.model_deflate_chunk.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        L.model_selfcheck.restype = C.c_int
        L.model_huffman_fuzz.restype = C.c_int
        L.model_huffman_fuzz.argtypes = [C.c_int, C.c_uint]
        _lib = L
    return _lib


def host_inflate(comp, cap, lbits=10, misalign=0):
    """Run inflate_core.h (G=1).  Returns (bytes, dict(status, produced, consumed, blocks))."""
    c = np.ascontiguousarray(comp, dtype=np.uint8)
    buf = np.full(cap + 64, 0xA5, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16 + misalign
    res = np.zeros(4, np.uint32)
    fn = lib().host_inflate_chunk
    fn(c.ctypes.data if c.size else None, c.size, base + off, cap, res.ctypes.data, lbits)
    out = buf[off:off + int(res[0])].copy()
    guard_ok = bool((buf[:off] == 0xA5).all() and (buf[off + cap:] == 0xA5).all())
    return out, {"produced": int(res[0]), "status": int(res[1]), "consumed": int(res[2]),
                 "blocks": int(res[3]), "guard_ok": guard_ok}


def host_inflate_fast(comp, cap, lbits=9, misalign=0, checksum_type=0):
    """Run the lane-per-stream decoder of the production inflate kernel (inflate_fast.h) with one lane."""
    c = np.ascontiguousarray(comp, dtype=np.uint8)
    buf = np.full(cap + 64, 0xA5, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16 + misalign
    res = np.zeros(6, np.uint32)
    lib().host_inflate_fast(c.ctypes.data if c.size else None, c.size, base + off, cap, res.ctypes.data, lbits, checksum_type)
    out = buf[off:off + int(res[0])].copy()
    guard_ok = bool((buf[:off] == 0xA5).all() and (buf[off + cap:] == 0xA5).all())
    return out, {"produced": int(res[0]), "status": int(res[1]), "consumed": int(res[2]), "blocks": int(res[3]),
                 "crc32": int(res[4]), "adler32": int(res[5]), "guard_ok": guard_ok}


def host_inflate_indexed(comp, cap, misalign=0):
    """The indexed (sub-range parallel) path of the inflate kernel, run sequentially on the CPU."""
    c = np.ascontiguousarray(comp, dtype=np.uint8)
    buf = np.full(cap + 64, 0xA5, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16 + misalign
    res = np.zeros(4, np.uint32)
    lib().host_inflate_indexed(c.ctypes.data if c.size else None, c.size, base + off, cap, res.ctypes.data)
    out = buf[off:off + int(res[0])].copy()
    guard_ok = bool((buf[:off] == 0xA5).all() and (buf[off + cap:] == 0xA5).all())
    return out, {"produced": int(res[0]), "status": int(res[1]), "indexed": bool(res[2]), "subs": int(res[3]),
                 "guard_ok": guard_ok}


INDEX_MAGIC = 0xB17A0B02
SUB = 2048


def split_index(stream):
    """Splits a chunk produced by the GPU deflate kernel / its model into (deflate bytes, index | None).
    The index is the parallel-inflate trailer described in bitar_b200/csrc/deflate_common.h; all its
    structural invariants are checked here."""
    s = np.ascontiguousarray(stream, dtype=np.uint8)
    if s.size < 12 or int(s[-4:].view("<u4")[0]) != INDEX_MAGIC:
        return s, None
    total = int(s[-8:-4].view("<u4")[0])
    end_bit = int(s[-12:-8].view("<u4")[0])
    full, rem = divmod(total, 65536)
    entries = full * 33 + ((1 + (rem + SUB - 1) // SUB) if rem else 0)
    nbytes = 4 * (entries + 3)
    if (end_bit + 7) // 8 + nbytes != s.size:
        return s, None
    words = s[s.size - nbytes:s.size - 12].view("<u4").astype(np.int64)
    blocks = []
    at = 0
    for b in range(full + (1 if rem else 0)):
        blen = 65536 if b < full else rem
        ns = (blen + SUB - 1) // SUB
        blocks.append({"hdr_bit": int(words[at]), "sub_bit": [int(x) for x in words[at + 1:at + 1 + ns]], "len": blen})
        at += 1 + ns
    flat = [v for b in blocks for v in [b["hdr_bit"]] + [x for x in b["sub_bit"] if x]]
    assert flat == sorted(flat) and (not flat or flat[-1] < end_bit), "index offsets must increase"
    return s[:s.size - nbytes], {"blocks": blocks, "end_bit": end_bit, "total_out": total}


def model_deflate(data, huffman=2, block=0):
    d = np.ascontiguousarray(data, dtype=np.uint8)
    cap = d.size + d.size // 8 + 1024
    out = np.empty(cap, np.uint8)
    r = lib().model_deflate_chunk(d.ctypes.data if d.size else None, d.size, out.ctypes.data, cap, huffman, block)
    if r < 0:
        raise RuntimeError("model deflate overflow")
    return out[:r].copy()


def host_inflate_spec(comp, cap, target=768, misalign=0):
    """Run the speculative lane-parallel decoder of streams without an index (inflate_spec.h), the 32 lanes of a round one
    after the other.  Returns (bytes, dict(produced, declined, blocks, rounds, ranges, short_rounds, walk_symbols))."""
    c = np.ascontiguousarray(comp, dtype=np.uint8)
    cpad = np.concatenate([c, np.zeros(8, np.uint8)])          # (the bit reader reads whole aligned words)
    buf = np.full(cap + 64, 0xA5, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16 + misalign
    res = np.zeros(12, np.uint32)
    lib().host_inflate_spec(cpad.ctypes.data, c.size, base + off, cap, res.ctypes.data, target)
    out = buf[off:off + int(res[0])].copy()
    guard_ok = bool((buf[:off] == 0xA5).all() and (buf[off + cap:] == 0xA5).all())
    return out, {"produced": int(res[0]), "declined": int(res[1]), "blocks": int(res[2]), "rounds": int(res[3]),
                 "ranges": int(res[4]), "short_rounds": int(res[5]), "walk_symbols": int(res[6]), "full_rounds": int(res[7]),
                 "warp_steps": int(res[8]), "lane_steps": int(res[9]), "guard_ok": guard_ok}
