"""ctypes loader for the CPU oracle (oracle/oracle.c) and, when it has been
built, the reference's own CPU implementation (oracle/_ref/libref_cpu.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(drjit-core_b200/) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_cpu.so")

# reference ABI: include/drjit-core/jit.h:597-611 and :990-1014
VT = dict(bool=1, i8=3, u8=4, i16=5, u16=6, i32=7, u32=8, i64=9, u64=10,
          f16=13, f32=14, f64=15)
OP = dict(add=1, mul=2, min=3, max=4, and_=5, or_=6)
NP_OF_VT = {4: np.uint8, 7: np.int32, 8: np.uint32, 9: np.int64, 10: np.uint64,
            13: np.float16, 14: np.float32, 15: np.float64}


def build(ref=True, quiet=True):
    """Compile liboracle.so (always) and oracle/_ref (only when the reference
    sources are present, i.e. in the build container)."""
    targets = ["liboracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE, "-j8"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self):
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        L = self.lib = ctypes.CDLL(_ORACLE_SO)
        L.oracle_fmix32.restype = ctypes.c_uint32
        L.oracle_fmix32.argtypes = [ctypes.c_uint32]
        L.oracle_reduce_identity.restype = ctypes.c_uint64
        L.oracle_reduce_identity.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_type_size.restype = ctypes.c_uint32
        L.oracle_compress.restype = ctypes.c_uint32
        L.oracle_block_mkperm.restype = ctypes.c_uint32
        L.oracle_half_to_float.restype = ctypes.c_float
        L.oracle_half_to_float.argtypes = [ctypes.c_uint16]
        L.oracle_float_to_half.restype = ctypes.c_uint16
        L.oracle_float_to_half.argtypes = [ctypes.c_float]

    def block_reduce(self, vt, op, x, block_size, wide=False):
        size = x.shape[0]
        blocks = (size + block_size - 1) // block_size if block_size else 0
        out = np.zeros(blocks, dtype=x.dtype)
        rc = self.lib.oracle_block_reduce(vt, op, size, block_size, _ptr(x),
                                          _ptr(out), int(wide))
        if rc:
            raise ValueError(f"oracle_block_reduce: rc={rc}")
        return out

    def block_prefix_reduce(self, vt, op, x, block_size, exclusive, reverse,
                            wide=False):
        out = np.zeros_like(x)
        rc = self.lib.oracle_block_prefix_reduce(
            vt, op, x.shape[0], block_size, int(exclusive), int(reverse),
            _ptr(x), _ptr(out), int(wide))
        if rc:
            raise ValueError(f"oracle_block_prefix_reduce: rc={rc}")
        return out

    def reduce_dot(self, vt, a, b, wide=False):
        out = np.zeros(1, dtype=a.dtype)
        rc = self.lib.oracle_reduce_dot(vt, _ptr(a), _ptr(b), a.shape[0],
                                        _ptr(out), int(wide))
        if rc:
            raise ValueError(f"oracle_reduce_dot: rc={rc}")
        return out

    def compress(self, mask):
        out = np.zeros(max(1, mask.shape[0]), dtype=np.uint32)
        n = self.lib.oracle_compress(_ptr(mask), mask.shape[0], _ptr(out))
        return out[:n].copy(), n

    def block_mkperm(self, keys, block_size, bucket_count, want_offsets=True):
        size = keys.shape[0]
        perm = np.zeros(max(1, size), dtype=np.uint32)
        offsets = (np.zeros(4 * bucket_count + 1, dtype=np.uint32)
                   if want_offsets else None)
        n = self.lib.oracle_block_mkperm(_ptr(keys), size, block_size,
                                         bucket_count, _ptr(perm), _ptr(offsets))
        if n == 0xFFFFFFFF:
            raise ValueError("oracle_block_mkperm: key out of range")
        return perm[:size], offsets, n

    def scatter_reduce(self, vt, op, target, value, index, mask=None, wide=False):
        tgt = target.copy()
        rc = self.lib.oracle_scatter_reduce(
            vt, op, _ptr(tgt), tgt.shape[0], _ptr(value), _ptr(index),
            _ptr(mask), index.shape[0], int(wide))
        if rc:
            raise ValueError(f"oracle_scatter_reduce: rc={rc}")
        return tgt

    def scatter_reduce_packet(self, vt, op, target, values, index, mask=None, wide=False):
        """src/cuda_packet.cpp:169-327 applied serially: target[index * W + k] op=
        values[k] -- W independent scatters on the strided views of the target."""
        width = len(values)
        tgt = target.copy().reshape(-1, width)
        for k in range(width):
            col = self.scatter_reduce(vt, op, np.ascontiguousarray(tgt[:, k]), values[k], index, mask, wide)
            tgt[:, k] = col
        return tgt.reshape(-1)

    @staticmethod
    def scatter_packet(target, values, index, mask=None):
        """src/cuda_packet.cpp:329-443 applied serially (element order): target[index * W
        + k] = values[k].  With duplicate indices the reference leaves it open which value
        wins; callers compare only inputs without active duplicates."""
        width = len(values)
        tgt = target.copy().reshape(-1, width)
        on = np.ones(index.shape[0], dtype=bool) if mask is None else mask.astype(bool)
        for k in range(width):
            tgt[index[on], k] = values[k][on]
        return tgt.reshape(-1)

    @staticmethod
    def gather_packet(source, width, index, mask=None):
        """src/cuda_packet.cpp:18-166: out[k][i] = source[index[i] * W + k], masked-off
        lanes read 0 (:52-56)."""
        src = source.reshape(-1, width)
        on = np.ones(index.shape[0], dtype=bool) if mask is None else mask.astype(bool)
        outs = []
        for k in range(width):
            o = np.zeros(index.shape[0], dtype=source.dtype)
            o[on] = src[index[on], k]
            outs.append(o)
        return outs

    @staticmethod
    def scatter_inc_check(target_before, target_after, index, mask, out):
        """Semantics of src/cuda_scatter.cpp:356-393 (the order in which entries of
        one counter are served is unspecified): every counter grows by the number
        of its active entries, those entries receive exactly the consecutive old
        values base .. base + count - 1, masked entries receive 0.  Returns a list
        of violations (empty = pass)."""
        on = np.ones(index.shape[0], dtype=bool) if mask is None else mask.astype(bool)
        bad = []
        cnt = np.bincount(index[on], minlength=target_before.shape[0]).astype(np.uint32)
        if not np.array_equal((target_before + cnt).astype(np.uint32), target_after):
            bad.append("counters")
        if np.any(out[~on] != 0):
            bad.append("masked entries not zero")
        order = np.lexsort((out[on], index[on]))
        si, so = index[on][order], out[on][order].astype(np.int64)
        start = np.r_[True, si[1:] != si[:-1]]
        rank = np.arange(si.size) - np.maximum.accumulate(np.where(start, np.arange(si.size), 0))
        if not np.array_equal(so, target_before[si].astype(np.int64) + rank):
            bad.append("old values are not base + 0..count-1")
        return bad

    def all(self, mask):
        return bool(self.lib.oracle_all(_ptr(mask), mask.shape[0]))

    def any(self, mask):
        return bool(self.lib.oracle_any(_ptr(mask), mask.shape[0]))

    def reduce_identity(self, vt, op):
        return self.lib.oracle_reduce_identity(vt, op)


def fmix32(i):
    """Vectorised tests/reductions.cpp:5-13 on a numpy uint32 array."""
    h = (np.asarray(i, dtype=np.uint32) + np.uint32(1)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85ebca6b)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xc2b2ae35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def ref_available():
    return os.path.exists(_REF_SO)


class Reference:
    """The reference's own CPU ("LLVM backend") implementation of the path,
    compiled from /root/reference by oracle/Makefile into oracle/_ref/."""

    def __init__(self, threads=0):
        if not ref_available():
            raise RuntimeError("oracle/_ref/libref_cpu.so has not been built")
        L = self.lib = ctypes.CDLL(_REF_SO)
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_pool_size.restype = ctypes.c_uint32
        L.ref_reduce_identity.restype = ctypes.c_uint64
        if L.ref_init(ctypes.c_uint32(threads)):
            raise RuntimeError(L.ref_last_error().decode())

    def _check(self, rc):
        if rc:
            raise ValueError(self.lib.ref_last_error().decode())

    @property
    def threads(self):
        return int(self.lib.ref_pool_size())

    def block_reduce(self, vt, op, x, block_size, out=None):
        size = x.shape[0]
        if out is None:
            blocks = (size + block_size - 1) // block_size if block_size else 0
            out = np.zeros(blocks, dtype=x.dtype)
        self._check(self.lib.ref_block_reduce(vt, op, size, block_size,
                                              _ptr(x), _ptr(out)))
        return out

    def block_prefix_reduce(self, vt, op, x, block_size, exclusive, reverse,
                            out=None):
        if out is None:
            out = np.zeros_like(x)
        self._check(self.lib.ref_block_prefix_reduce(
            vt, op, x.shape[0], block_size, int(exclusive), int(reverse),
            _ptr(x), _ptr(out)))
        return out

    def reduce_dot(self, vt, a, b):
        out = np.zeros(1, dtype=a.dtype)
        self._check(self.lib.ref_reduce_dot(vt, _ptr(a), _ptr(b), a.shape[0],
                                            _ptr(out)))
        return out

    def compress(self, mask, out=None):
        if out is None:
            out = np.zeros(max(1, mask.shape[0]), dtype=np.uint32)
        cnt = ctypes.c_uint32(0)
        self._check(self.lib.ref_compress(_ptr(mask), mask.shape[0], _ptr(out),
                                          ctypes.byref(cnt)))
        return out[:cnt.value], cnt.value

    def block_mkperm(self, keys, block_size, bucket_count, want_offsets=True,
                     perm=None):
        size = keys.shape[0]
        if perm is None:
            perm = np.zeros(max(1, size), dtype=np.uint32)
        offsets = (np.zeros(4 * bucket_count + 1, dtype=np.uint32)
                   if want_offsets else None)
        uq = ctypes.c_uint32(0)
        self._check(self.lib.ref_block_mkperm(
            _ptr(keys), size, block_size, bucket_count, _ptr(perm),
            _ptr(offsets), ctypes.byref(uq)))
        return perm[:size], offsets, uq.value

    def all(self, mask):
        buf = np.zeros(mask.shape[0] + 4, dtype=np.uint8)  # reference pads in place
        buf[:mask.shape[0]] = mask
        r = ctypes.c_int(0)
        self._check(self.lib.ref_all(_ptr(buf), mask.shape[0], ctypes.byref(r)))
        return bool(r.value)

    def any(self, mask):
        buf = np.zeros(mask.shape[0] + 4, dtype=np.uint8)
        buf[:mask.shape[0]] = mask
        r = ctypes.c_int(0)
        self._check(self.lib.ref_any(_ptr(buf), mask.shape[0], ctypes.byref(r)))
        return bool(r.value)

    def reduce_identity(self, vt, op):
        return self.lib.ref_reduce_identity(vt, op)


_REF_CUDA_SO = os.path.join(_HERE, "_ref", "libref_cuda.so")


def ref_cuda_available():
    return os.path.exists(_REF_CUDA_SO)


class ReferenceCUDA:
    """The reference built WITH its CUDA backend (oracle/_ref/libref_cuda.so):
    its own CUDA kernels (compute_75 PTX, JIT-compiled by the driver) driven
    through the public jit_* entry points on raw DEVICE pointers (ints, e.g.
    torch's ``tensor.data_ptr()``).  GPU box only.  All calls are enqueued on
    the reference's own stream; ``sync()`` waits for it."""

    _instance = None

    @classmethod
    def get(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        if not ref_cuda_available():
            raise RuntimeError("oracle/_ref/libref_cuda.so has not been built")
        L = self.lib = ctypes.CDLL(_REF_CUDA_SO)
        L.refcuda_last_error.restype = ctypes.c_char_p
        L.refcuda_stream.restype = ctypes.c_void_p
        vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
        L.refcuda_block_reduce.argtypes = [i32, i32, u32, u32, vp, vp]
        L.refcuda_block_prefix_reduce.argtypes = [i32, i32, u32, u32, i32, i32, vp, vp]
        L.refcuda_reduce_dot.argtypes = [i32, vp, vp, u32, vp]
        L.refcuda_compress.argtypes = [vp, u32, vp, ctypes.POINTER(u32)]
        L.refcuda_block_mkperm.argtypes = [vp, u32, u32, u32, vp, vp, ctypes.POINTER(u32)]
        L.refcuda_scatter_reduce.argtypes = [i32, i32, i32, vp, ctypes.c_size_t, vp, vp, vp,
                                             ctypes.c_size_t, i32]
        L.refcuda_all.argtypes = [vp, u32, ctypes.POINTER(i32)]
        L.refcuda_any.argtypes = [vp, u32, ctypes.POINTER(i32)]
        L.refcuda_can_scatter_reduce.argtypes = [i32, i32]
        L.refcuda_scatter_inc.argtypes = [vp, ctypes.c_size_t, vp, vp, ctypes.c_size_t, vp]
        L.refcuda_scatter_packet.argtypes = [i32, i32, i32, vp, ctypes.c_size_t, ctypes.POINTER(vp),
                                             ctypes.c_size_t, vp, vp, ctypes.c_size_t]
        if L.refcuda_init():
            raise RuntimeError(L.refcuda_last_error().decode())

    def _check(self, rc):
        if rc:
            raise ValueError(self.lib.refcuda_last_error().decode())

    def sync(self):
        self.lib.refcuda_sync()

    def block_reduce(self, vt, op, size, block_size, d_in, d_out):
        self._check(self.lib.refcuda_block_reduce(vt, op, size, block_size, d_in, d_out))

    def block_prefix_reduce(self, vt, op, size, block_size, exclusive, reverse, d_in, d_out):
        self._check(self.lib.refcuda_block_prefix_reduce(vt, op, size, block_size, int(exclusive),
                                                         int(reverse), d_in, d_out))

    def reduce_dot(self, vt, d_a, d_b, size, d_out):
        self._check(self.lib.refcuda_reduce_dot(vt, d_a, d_b, size, d_out))

    def compress(self, d_mask, size, d_out):
        """NOTE: the reference zero-fills the mask buffer past 'size' (up to the
        next power of two >= size, src/cuda_ts.cpp:708-710,746-748)."""
        cnt = ctypes.c_uint32(0)
        self._check(self.lib.refcuda_compress(d_mask, size, d_out, ctypes.byref(cnt)))
        return cnt.value

    def block_mkperm(self, d_keys, size, block_size, bucket_count, d_perm, h_offsets_pinned):
        uq = ctypes.c_uint32(0)
        self._check(self.lib.refcuda_block_mkperm(d_keys, size, block_size, bucket_count, d_perm,
                                                  h_offsets_pinned, ctypes.byref(uq)))
        return uq.value

    def scatter_reduce(self, vt, op, mode, d_target, target_size, d_value, d_index, d_mask, n,
                       repeat=1):
        self._check(self.lib.refcuda_scatter_reduce(vt, op, mode, d_target, target_size, d_value,
                                                    d_index, d_mask, n, repeat))

    def can_scatter_reduce(self, vt, op):
        return bool(self.lib.refcuda_can_scatter_reduce(vt, op))

    def scatter_inc(self, d_target, target_size, d_index, d_mask, n, d_out):
        self._check(self.lib.refcuda_scatter_inc(d_target, target_size, d_index, d_mask, n, d_out))

    def scatter_packet(self, vt, op, mode, d_target, target_size, d_values, d_index, d_mask, n):
        ptrs = (ctypes.c_void_p * len(d_values))(*d_values)
        self._check(self.lib.refcuda_scatter_packet(vt, op, mode, d_target, target_size, ptrs,
                                                    len(d_values), d_index, d_mask, n))

    def all(self, d_mask, size):
        r = ctypes.c_int(0)
        self._check(self.lib.refcuda_all(d_mask, size, ctypes.byref(r)))
        return bool(r.value)

    def any(self, d_mask, size):
        r = ctypes.c_int(0)
        self._check(self.lib.refcuda_any(d_mask, size, ctypes.byref(r)))
        return bool(r.value)
