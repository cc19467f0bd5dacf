"""Helpers for the GPU parity tests: device buffers through torch, batch (de)compression of chunk
lists through the C-ABI."""
import numpy as np
import torch

from bitar_b200 import _capi as capi
from bitar_b200.engine import CompressDevice, Configuration


def to_dev(arr, pad=64):
    """numpy uint8 -> cuda uint8 tensor (with tail padding so neighbouring reads stay in-bounds)."""
    t = torch.zeros(arr.size + pad, dtype=torch.uint8, device="cuda")
    if arr.size:
        t[:arr.size] = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    torch.cuda.synchronize()
    return t


def open_device(seg=59460, qps=1, **kw):
    cfg = Configuration(decompressed_seg_size=seg, **kw)
    return CompressDevice(0, qps).Initialize(cfg)


def pack_chunks(chunks, align=16, shift=0):
    """Concatenate byte chunks at `align`-aligned (+shift) offsets. Returns (buffer, offsets)."""
    offs, at = [], 0
    for c in chunks:
        at = (at + align - 1) // align * align + shift
        offs.append(at)
        at += c.size
    buf = np.zeros(at + 64, np.uint8)
    for c, o in zip(chunks, offs):
        buf[o:o + c.size] = c
    return buf, np.array(offs, np.uint64)


def gpu_deflate_chunks(dev, chunks, qp=0, src_shift=0, dst_shift=0, cap=None):
    """Deflate every chunk as one op. Returns (list of compressed np arrays, results)."""
    n = len(chunks)
    buf, offs = pack_chunks(chunks, 16, src_shift)
    src = to_dev(buf)
    cap_each = cap if cap is not None else max([capi.lib().bitar_compressed_seg_size(max(c.size, 8)) for c in chunks] + [64])
    stride = (cap_each + 15) // 16 * 16 + 16
    dst = torch.full((n * stride + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ops = np.zeros(n, capi.CHUNK_DTYPE)
    ops["src"] = np.uint64(src.data_ptr()) + offs
    ops["src_len"] = [c.size for c in chunks]
    ops["dst"] = np.uint64(dst.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(stride) + np.uint64(dst_shift)
    ops["dst_cap"] = cap_each
    res = dev.enqueue("deflate", qp, ops)
    try:
        dev.wait(qp)
        err = None
    except capi.BitarError as e:
        err = e
    host = dst.cpu().numpy()
    outs = []
    for i in range(n):
        o = i * stride + dst_shift
        outs.append(host[o:o + int(res["produced"][i])].copy())
        # bytes outside [o, o+cap) must be untouched
        assert (host[i * stride:o] == 0xA5).all(), "deflate wrote before dst"
        assert (host[o + cap_each:(i + 1) * stride] == 0xA5).all(), "deflate wrote past dst_cap"
    return outs, res, err


def gpu_inflate_chunks(dev, comps, caps, qp=0, src_shift=0, dst_shift=0):
    """Inflate every compressed chunk as one op. Returns (list of np arrays, results, error)."""
    n = len(comps)
    buf, offs = pack_chunks(comps, 4, src_shift)
    src = to_dev(buf)
    stride = (max(list(caps) + [16]) + 15) // 16 * 16 + 32
    dst = torch.full((n * stride + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ops = np.zeros(n, capi.CHUNK_DTYPE)
    ops["src"] = np.uint64(src.data_ptr()) + offs
    ops["src_len"] = [c.size for c in comps]
    ops["dst"] = np.uint64(dst.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(stride) + np.uint64(dst_shift)
    ops["dst_cap"] = caps
    res = dev.enqueue("inflate", qp, ops)
    try:
        dev.wait(qp)
        err = None
    except capi.BitarError as e:
        err = e
    host = dst.cpu().numpy()
    outs = []
    for i in range(n):
        o = i * stride + dst_shift
        outs.append(host[o:o + int(res["produced"][i])].copy())
        assert (host[i * stride:o] == 0xA5).all(), "inflate wrote before dst"
        assert (host[o + int(caps[i]):(i + 1) * stride] == 0xA5).all(), "inflate wrote past dst_cap"
    return outs, res, err
