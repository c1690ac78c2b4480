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
