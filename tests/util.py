"""Helpers for the GPU parity tests: numpy <-> device transfer that keeps
unsigned and half types bit-exact."""
import hashlib

import numpy as np

_VIEW = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64,
         np.dtype(np.uint16): np.int16}


def to_dev(a, offset_elems=0):
    """numpy -> CUDA tensor (optionally placed 'offset_elems' elements into a
    larger allocation to produce misaligned pointers)."""
    import torch
    a = np.ascontiguousarray(a)
    v = a.view(_VIEW[a.dtype]) if a.dtype in _VIEW else a
    t = torch.from_numpy(v)
    if offset_elems:
        buf = torch.empty(t.numel() + offset_elems + 64, dtype=t.dtype, device="cuda")
        out = buf[offset_elems:offset_elems + t.numel()]
        out.copy_(t)
        return out
    return t.cuda()


def empty_dev(n, dtype, offset_elems=0):
    import torch
    dtype = np.dtype(dtype)
    vd = np.dtype(_VIEW.get(dtype, dtype))
    tdt = getattr(torch, vd.name)
    buf = torch.zeros(max(n, 1) + offset_elems + 64, dtype=tdt, device="cuda")
    return buf[offset_elems:offset_elems + n]


def to_host(t, dtype):
    import torch
    torch.cuda.synchronize()
    a = t.detach().cpu().numpy()
    return a.view(dtype)


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-30)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0


# ---- counter-based inputs generated ON THE DEVICE (same values as tests/golden/cases.py)
def dev_fmix32(n, xor=0, index=None, start=0):
    """fmix32(i ^ xor) for i in [start, start + n) as an int32 CUDA tensor holding the
    uint32 bit pattern (tests/reductions.cpp:5-13 of the reference)."""
    import torch
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    chunk = 1 << 26
    M = 0xFFFFFFFF
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        if index is not None:
            i = index[s:e].to(torch.int64)
        else:
            i = torch.arange(start + s, start + e, dtype=torch.int64, device="cuda")
        if xor:
            i = i ^ xor
        h = ((i & M) + 1) & M
        h = h ^ (h >> 16)
        h = (h * 0x85ebca6b) & M
        h = h ^ (h >> 13)
        h = (h * 0xc2b2ae35) & M
        h = h ^ (h >> 16)
        out[s:e] = (h - ((h >> 31) << 32)).to(torch.int32)
    return out


def dev_u32_input(n, start=0):
    return dev_fmix32(n, start=start)


def dev_f32_input(n, start=0):
    """uniform [0, 1) with 24 random bits (cases.f32_input)"""
    import torch
    h = dev_fmix32(n, start=start)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    chunk = 1 << 26
    for s in range(0, n, chunk):
        hu = h[s:s + chunk].to(torch.int64) & 0xFFFFFFFF
        out[s:s + chunk] = (hu >> 8).to(torch.float32) * (2.0 ** -24)
    return out
