"""ctypes binding of the C-ABI in include/bitar_cuda.h (libbitar_cuda.so).

This is the only way the Python side reaches the engine: there is no CPU fallback.  Loading fails
loudly when the library has not been built (run ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C bitar_b200/csrc``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BITAR_CUDA_LIB") or os.path.join(_HERE, "csrc", "libbitar_cuda.so")   # env: A/B runs of two builds

# struct bitar_chunk / bitar_result as numpy record dtypes (layout-identical to the C structs)
CHUNK_DTYPE = np.dtype([("src", "<u8"), ("dst", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4")])
RESULT_DTYPE = np.dtype([("produced", "<u4"), ("status", "<u4"), ("checksum", "<u8")])
assert CHUNK_DTYPE.itemsize == 24 and RESULT_DTYPE.itemsize == 16

OK, E_OUT_OF_MEMORY, E_INVALID, E_IO_ERROR, E_CAPACITY, E_CANCELLED, E_UNKNOWN, E_NOT_IMPLEMENTED = 0, -1, -4, -5, -6, -8, -9, -10
OP_OK, OP_OUT_OF_SPACE, OP_DATA_ERROR, OP_TRUNCATED = 0, 1, 2, 3
HUFFMAN_DEFAULT, HUFFMAN_FIXED, HUFFMAN_DYNAMIC = 0, 1, 2
CHECKSUM_NONE, CHECKSUM_CRC32, CHECKSUM_ADLER32, CHECKSUM_CRC32_ADLER32 = 0, 1, 2, 3
MEM_DEVICE, MEM_PINNED = 0, 1

STATUS_NAME = {0: "OK", -1: "OutOfMemory", -4: "Invalid", -5: "IOError", -6: "CapacityError",
               -8: "Cancelled", -9: "UnknownError", -10: "NotImplemented"}


class DevInfo(C.Structure):
    _fields_ = [("device_id", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("total_mem", C.c_uint64), ("max_queue_pairs", C.c_uint32),
                ("window_min", C.c_uint8), ("window_max", C.c_uint8),
                ("supports_fixed", C.c_uint8), ("supports_dynamic", C.c_uint8),
                ("supports_crc32", C.c_uint8), ("supports_adler32", C.c_uint8),
                ("supports_sgl", C.c_uint8), ("reserved", C.c_uint8), ("name", C.c_char * 64)]


class Cfg(C.Structure):
    _fields_ = [("decompressed_seg_size", C.c_uint32), ("compressed_seg_size", C.c_uint32),
                ("max_preallocate_slots", C.c_uint32), ("burst_size", C.c_uint16),
                ("max_sgl_segs", C.c_uint16), ("window_size", C.c_uint8), ("huffman_enc", C.c_uint8),
                ("checksum_type", C.c_uint8), ("slot_mem_kind", C.c_uint8), ("no_index", C.c_uint8),
                ("reserved", C.c_uint8 * 3)]


# every symbol include/bitar_cuda.h declares (tests check that the library exports all of them)
EXPORTS = [
    "bitar_cuda_device_count", "bitar_cuda_device_info", "bitar_compressed_seg_size",
    "bitar_reference_compressed_seg_size", "bitar_dev_open", "bitar_dev_close", "bitar_dev_config",
    "bitar_dev_num_qps", "bitar_qp_deflate", "bitar_qp_inflate", "bitar_qp_wait", "bitar_qp_result", "bitar_qp_busy",
    "bitar_qp_on_complete", "bitar_qp_last_ms", "bitar_qp_stream", "bitar_kernel_launches",
    "bitar_slot_take", "bitar_slot_take_n", "bitar_slot_put", "bitar_slot_put_n", "bitar_slot_size", "bitar_slots_free",
    "bitar_mem_alloc", "bitar_mem_free", "bitar_host_register", "bitar_host_unregister", "bitar_ptr_kind", "bitar_mem_copy", "bitar_current_device", "bitar_set_device",
    "bitar_qp_memcpy", "bitar_last_error", "bitar_version",
]

_lib = None


class BitarError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS_NAME.get(code, code)}: {msg}")
        self.code = code


def lib():
    """Load libbitar_cuda.so; raises if it was not built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first "
                          "(make -C bitar_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u16, u32, i32 = C.c_void_p, C.c_uint16, C.c_uint32, C.c_int
    L.bitar_cuda_device_count.restype = i32
    L.bitar_cuda_device_info.argtypes = [i32, C.POINTER(DevInfo)]
    L.bitar_compressed_seg_size.restype = u32
    L.bitar_compressed_seg_size.argtypes = [u32]
    L.bitar_reference_compressed_seg_size.restype = u32
    L.bitar_reference_compressed_seg_size.argtypes = [u32]
    L.bitar_dev_open.argtypes = [i32, u16, C.POINTER(Cfg), C.POINTER(vp)]
    L.bitar_dev_close.argtypes = [vp]
    L.bitar_dev_config.argtypes = [vp, C.POINTER(Cfg)]
    L.bitar_dev_num_qps.restype = u16
    L.bitar_dev_num_qps.argtypes = [vp]
    for f in (L.bitar_qp_deflate, L.bitar_qp_inflate):
        f.argtypes = [vp, u16, vp, u32, vp]
    L.bitar_qp_wait.argtypes = [vp, u16]
    L.bitar_qp_busy.argtypes = [vp, u16]
    L.bitar_qp_result.argtypes = [vp, u16]
    L.bitar_qp_on_complete.argtypes = [vp, u16, vp, vp]
    L.bitar_qp_last_ms.argtypes = [vp, u16, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.bitar_qp_stream.restype = vp
    L.bitar_qp_stream.argtypes = [vp, u16]
    L.bitar_kernel_launches.restype = C.c_uint64
    L.bitar_slot_take.restype = vp
    L.bitar_slot_take.argtypes = [vp]
    L.bitar_slot_take_n.argtypes = [vp, u32, vp]
    L.bitar_slot_put.argtypes = [vp, vp]
    L.bitar_slot_put_n.restype = u32
    L.bitar_slot_put_n.argtypes = [vp, vp, u32]
    L.bitar_slot_size.restype = u32
    L.bitar_slot_size.argtypes = [vp]
    L.bitar_slots_free.restype = u32
    L.bitar_slots_free.argtypes = [vp]
    L.bitar_mem_alloc.argtypes = [i32, i32, C.c_size_t, C.c_size_t, C.POINTER(vp)]
    L.bitar_mem_free.argtypes = [i32, i32, vp]
    L.bitar_host_register.argtypes = [vp, C.c_size_t]
    L.bitar_host_unregister.argtypes = [vp]
    L.bitar_ptr_kind.argtypes = [vp, C.POINTER(C.c_int)]
    L.bitar_mem_copy.argtypes = [vp, vp, C.c_size_t]
    L.bitar_current_device.argtypes = [C.POINTER(C.c_int)]
    L.bitar_qp_memcpy.argtypes = [vp, u16, vp, vp, C.c_size_t]
    L.bitar_last_error.restype = C.c_char_p
    L.bitar_version.restype = C.c_char_p
    L.bitar_tune_inflate_variant.argtypes = [i32]
    if hasattr(L, "bitar_tune_spec_target"):   # (test hooks; absent from older builds loaded through BITAR_CUDA_LIB for A/B runs)
        L.bitar_tune_spec_target.argtypes = [i32]
        L.bitar_debug_inflate_counters.argtypes = [vp, u16, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise BitarError(rc, lib().bitar_last_error().decode())
    return rc
