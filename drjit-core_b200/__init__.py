"""drjit-core_b200 -- B200-native (sm_100a) data-parallel primitives behind the
drjit-core `jit.h` entry points: block/whole-array reductions, dot products,
blocked prefix reductions, mask compression, the bucketing permutation of
vectorised method dispatch (block_mkperm) and atomic scatter-reduce.

This module is a thin ctypes binding of the C-ABI declared in
include/drjit_b200.h (libdrjit_core_b200.so, built in-tree from csrc/*.cu).
Function names, argument order, enum values and error behaviour follow the
reference (include/drjit-core/jit.h); pointers may be given as integers or as
torch tensors (their data_ptr() is used).  There is NO CPU fallback: if the
CUDA library is missing, or no GPU is present, calls raise.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# (B200_LIB_PATH: development builds, e.g. the scan tuning build of tools/tune_scan.py)
_LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "libdrjit_core_b200.so")


# ---- reference enums (jit.h:47-61, :597-611, :990-1014, :1017-1066) ----------
class JitBackend:
    None_, CUDA, LLVM, Metal = 0, 1, 2, 3


class VarType:
    (Void, Bool, BaseInt, Int8, UInt8, Int16, UInt16, Int32, UInt32, Int64, UInt64,
     Pointer, BaseFloat, Float16, Float32, Float64) = range(16)


class ReduceOp:
    Identity, Add, Mul, Min, Max, And, Or = range(7)


class ReduceMode:
    Auto, Direct, Local, NoConflicts, Expand, Permute = range(6)


class JitFlag:  # jit.h:1734-1742 (the bits this path honours)
    KernelHistory, LaunchBlocking, ForbidSynchronization = 1 << 15, 1 << 16, 1 << 17


class KernelType:  # jit.h:2597-2634
    (JIT, BlockReduce, BlockPrefixReduce, Dot, BatchedGemm, Compress, MkPerm, Memcpy, Memset,
     Poke, Aggregate, LLVMHostFunc) = range(12)


class KernelRecord(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("size", ctypes.c_uint64),
                ("execution_time_ms", ctypes.c_float)]


TYPE_SIZE = (0, 1, 0, 1, 1, 2, 2, 4, 4, 8, 8, 8, 0, 2, 4, 8)  # src/var.cpp:117-119


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libdrjit_core_b200.so (in-tree)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None
_LEGACY_STREAM = 1  # cudaStreamLegacy


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run drjit-core_b200.build() / "
                "__graft_entry__.build() first. There is no CPU fallback.")
        L = ctypes.CDLL(_LIB_PATH)
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_stream.restype = ctypes.c_void_p
        L.b200_malloc.restype = ctypes.c_void_p
        L.b200_malloc.argtypes = [ctypes.c_size_t, ctypes.c_int]
        L.b200_free.argtypes = [ctypes.c_void_p]
        L.b200_free_on.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.b200_malloc_migrate.restype = ctypes.c_void_p
        L.b200_malloc_migrate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.b200_sync_stream.argtypes = [ctypes.c_void_p]
        L.b200_set_flags.argtypes = [ctypes.c_uint32]
        L.b200_set_flags.restype = None
        L.b200_flags.restype = ctypes.c_uint32
        L.b200_set_flag.argtypes = [ctypes.c_uint32, ctypes.c_int]
        L.b200_set_flag.restype = None
        L.b200_kernel_history.argtypes = [ctypes.POINTER(KernelRecord), ctypes.c_int]
        L.b200_kernel_history_clear.restype = None
        L.b200_reduce_identity.restype = ctypes.c_uint64
        L.b200_launch_count.restype = ctypes.c_uint64
        vp, u64, u32, i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.b200_sync.argtypes = [vp]
        L.b200_memcpy.argtypes = [vp, vp, ctypes.c_size_t]
        L.b200_memcpy_async.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.b200_memset_async.argtypes = [vp, vp, u64, u32, vp]
        L.b200_block_reduce.argtypes = [vp, i, i, u64, u64, vp, vp]
        L.b200_reduce.argtypes = [vp, i, i, vp, u64, vp]
        L.b200_reduce_dot.argtypes = [vp, i, vp, vp, u64, vp]
        L.b200_block_prefix_reduce.argtypes = [vp, i, i, u64, u64, i, i, vp, vp]
        L.b200_prefix_reduce_carry.argtypes = [vp, i, i, u64, i, i, vp, vp, vp, vp]
        L.b200_prefix_reduce_seeded.argtypes = [vp, i, i, u64, i, i, vp, vp, vp]
        L.b200_scan_tile_elems.argtypes = [i]
        L.b200_scan_tile_elems.restype = u32
        L.b200_compress.argtypes = [vp, vp, u64, vp, ctypes.POINTER(u32)]
        L.b200_compress_async.argtypes = [vp, vp, u64, vp, vp]
        L.b200_block_mkperm.argtypes = [vp, vp, u32, u32, u32, vp, vp, ctypes.POINTER(u32)]
        L.b200_block_mkperm_async.argtypes = [vp, vp, u32, u32, u32, vp, vp]
        L.b200_mkperm_histogram.argtypes = [vp, vp, u64, u32, vp]
        L.b200_scatter_reduce.argtypes = [vp, i, i, i, vp, vp, vp, vp, u64]
        L.b200_scatter_inc.argtypes = [vp, vp, vp, vp, vp, u64]
        L.b200_scatter_reduce_packet.argtypes = [vp, i, i, i, vp, ctypes.POINTER(vp), u32, vp, vp, u64]
        L.b200_call_reduce.argtypes = [vp, vp, u32, u32, vp, vp, ctypes.POINTER(u32)]
        L.b200_call_reduce_async.argtypes = [vp, vp, u32, u32, vp, vp]
        L.b200_scatter_reduce_idx.argtypes = [vp, i, i, i, vp, vp, vp, i, vp, u64]
        L.b200_scatter_packet.argtypes = [vp, i, vp, ctypes.POINTER(vp), u32, vp, vp, u64]
        L.b200_gather_packet.argtypes = [vp, i, vp, ctypes.POINTER(vp), u32, vp, vp, u64]
        L.b200_sharded_create.argtypes = [i, i, ctypes.POINTER(vp)]
        L.b200_sharded_export.argtypes = [vp, vp]
        L.b200_sharded_connect.argtypes = [vp, vp]
        L.b200_sharded_destroy.argtypes = [vp]
        L.b200_sharded_reduce.argtypes = [vp, vp, i, i, vp, u64, vp]
        L.b200_sharded_reduce_dot.argtypes = [vp, vp, i, vp, vp, u64, vp]
        L.b200_sharded_prefix_reduce.argtypes = [vp, vp, i, i, u64, i, i, vp, vp]
        L.b200_sharded_prefix_reduce_cyclic.argtypes = [vp, vp, i, i, u64, u64, i, vp, vp]
        L.b200_sharded_histogram.argtypes = [vp, vp, vp, u64, u32, vp, vp]
        L.b200_all_async.argtypes = [vp, vp, u64, vp]
        L.b200_any_async.argtypes = [vp, vp, u64, vp]
        L.b200_all.argtypes = [vp, vp, u64, ctypes.POINTER(i)]
        L.b200_any.argtypes = [vp, vp, u64, ctypes.POINTER(i)]
        _lib = L
    return _lib


def _check(rc):
    if rc:
        # the reference raises std::runtime_error (src/log.cpp:165-169)
        raise RuntimeError(lib().b200_last_error().decode())


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):  # numpy (host memory: pinned buffers, offsets)
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


def _stream(stream):
    """None -> torch's current stream when torch is imported and CUDA is up,
    otherwise the library's own per-device stream."""
    if stream is not None:
        handle = stream if isinstance(stream, int) else stream.cuda_stream
        return handle or _LEGACY_STREAM
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        # torch's default stream is the legacy default stream, whose handle is 0;
        # the C-ABI reserves NULL for the library's own stream
        return torch.cuda.current_stream().cuda_stream or _LEGACY_STREAM
    return None


def _require_cuda(backend, name):
    if backend != JitBackend.CUDA:
        raise RuntimeError(f"{name}(): this build only provides the CUDA backend "
                           "(no CPU fallback).")


# ---- runtime -------------------------------------------------------------------
def jit_init(backends=1 << JitBackend.CUDA):
    if backends & (1 << JitBackend.CUDA):
        _check(lib().b200_init())


def jit_shutdown(light=0):
    _check(lib().b200_shutdown())


def jit_has_backend(backend):
    return backend == JitBackend.CUDA and lib().b200_init() == 0


def jit_sync_thread(stream=None):
    _check(lib().b200_sync(_stream(stream)))


def jit_cuda_device_count():
    return lib().b200_device_count()


def jit_cuda_set_device(device):
    _check(lib().b200_set_device(device))


def jit_cuda_stream():
    return lib().b200_stream()


def jit_malloc(backend, size, shared=0):
    _require_cuda(backend, "jit_malloc")
    ptr = lib().b200_malloc(size, 1 if shared else 0)
    if size and not ptr:
        _check(1)
    return ptr


def jit_free(ptr, stream=None):
    """Released in the order of the stream the binding launches on (torch's current
    stream when torch is up) -- the stream that last used the block."""
    s = _stream(stream)
    _check(lib().b200_free_on(s, ptr) if s else lib().b200_free(ptr))


def jit_malloc_migrate(ptr, backend, move=1):
    """jit.h:516: JitBackend.CUDA -> device memory, JitBackend.None_ -> pinned host."""
    if backend not in (JitBackend.CUDA, JitBackend.None_):
        _require_cuda(backend, "jit_malloc_migrate")
    r = lib().b200_malloc_migrate(ptr, 0 if backend == JitBackend.CUDA else 1, int(bool(move)))
    if not r:
        _check(1)
    return r


def jit_cuda_sync_stream(stream):
    """jit.h:243-255: `stream` (handle) waits for the work enqueued on the library stream."""
    _check(lib().b200_sync_stream(stream))


def jit_set_flags(flags):
    lib().b200_set_flags(flags)


def jit_flags():
    return int(lib().b200_flags())


def jit_set_flag(flag, enable):
    lib().b200_set_flag(flag, int(bool(enable)))


def jit_flag(flag):
    return bool(jit_flags() & flag)


def jit_kernel_history():
    """[(KernelType, size, execution_time_ms)] of the primitive calls recorded while
    JitFlag.KernelHistory was set; clears the history (jit.h:2715-2737)."""
    buf = (KernelRecord * 4096)()
    n = min(lib().b200_kernel_history(buf, 4096), 4096)
    return [(buf[i].type, int(buf[i].size), float(buf[i].execution_time_ms)) for i in range(n)]


def jit_kernel_history_clear():
    lib().b200_kernel_history_clear()


def jit_memcpy(backend, dst, src, size):
    _require_cuda(backend, "jit_memcpy")
    _check(lib().b200_memcpy(_ptr(dst), _ptr(src), size))


def jit_memcpy_async(backend, dst, src, size, stream=None):
    _require_cuda(backend, "jit_memcpy_async")
    _check(lib().b200_memcpy_async(_stream(stream), _ptr(dst), _ptr(src), size))


def jit_memset_async(backend, ptr, size, isize, src, stream=None):
    """src: bytes-like holding one element of isize bytes (jit.h:2199)."""
    _require_cuda(backend, "jit_memset_async")
    buf = (ctypes.c_uint8 * 8)(*bytes(src)[:8].ljust(8, b"\0"))
    _check(lib().b200_memset_async(_stream(stream), _ptr(ptr), size, isize, buf))


def jit_reduce_identity(vt, op):
    return lib().b200_reduce_identity(vt, op)


def launch_count():
    return int(lib().b200_launch_count())


def sm_count():
    return int(lib().b200_sm_count())


# ---- primitives (argument order of jit.h) ------------------------------------
def jit_block_reduce(backend, vt, op, size, block_size, in_, out, stream=None):
    _require_cuda(backend, "jit_block_reduce")
    _check(lib().b200_block_reduce(_stream(stream), vt, op, size, block_size,
                                   _ptr(in_), _ptr(out)))


def jit_reduce(backend, vt, op, in_, size, out, stream=None):
    _require_cuda(backend, "jit_reduce")
    _check(lib().b200_reduce(_stream(stream), vt, op, _ptr(in_), size, _ptr(out)))


def jit_reduce_dot(backend, vt, ptr_1, ptr_2, size, out, stream=None):
    _require_cuda(backend, "jit_reduce_dot")
    _check(lib().b200_reduce_dot(_stream(stream), vt, _ptr(ptr_1), _ptr(ptr_2), size,
                                 _ptr(out)))


def jit_block_prefix_reduce(backend, vt, op, size, block_size, exclusive, reverse,
                            in_, out, stream=None):
    """Positional contract of the reference: (.., size, block_size, ..)
    (src/api.cpp:1331-1337 -> src/util.cpp:55-61)."""
    _require_cuda(backend, "jit_block_prefix_reduce")
    _check(lib().b200_block_prefix_reduce(_stream(stream), vt, op, size, block_size,
                                          int(bool(exclusive)), int(bool(reverse)),
                                          _ptr(in_), _ptr(out)))


def prefix_reduce_carry(vt, op, size, exclusive, reverse, in_, out, carry_in=None,
                        carry_out=None, stream=None):
    _check(lib().b200_prefix_reduce_carry(_stream(stream), vt, op, size,
                                          int(bool(exclusive)), int(bool(reverse)),
                                          _ptr(in_), _ptr(out), _ptr(carry_in),
                                          _ptr(carry_out)))


def scan_tile_elems(vt):
    """Elements per tile of the streaming scan kernels (32 KiB)."""
    return lib().b200_scan_tile_elems(vt)


def prefix_reduce_seeded(vt, op, size, exclusive, reverse, in_, out, tile_seeds, stream=None):
    """Whole-array prefix reduction with caller-supplied exclusive tile prefixes."""
    _check(lib().b200_prefix_reduce_seeded(_stream(stream), vt, op, size, int(bool(exclusive)),
                                           int(bool(reverse)), _ptr(in_), _ptr(out),
                                           _ptr(tile_seeds)))


def jit_compress(backend, in_, size, out, stream=None):
    """Returns the number of non-zero mask entries (synchronises)."""
    _require_cuda(backend, "jit_compress")
    count = ctypes.c_uint32(0)
    _check(lib().b200_compress(_stream(stream), _ptr(in_), size, _ptr(out),
                               ctypes.byref(count)))
    return count.value


def compress_async(in_, size, out, count_dev, stream=None):
    _check(lib().b200_compress_async(_stream(stream), _ptr(in_), size, _ptr(out),
                                     _ptr(count_dev)))


def jit_block_mkperm(backend, values, size, block_size, bucket_count, perm, offsets,
                     stream=None):
    """offsets: host-accessible buffer of 4 * bucket_count + 1 uint32 (or None).
    Returns the number of unique values (0 if offsets is None / several groups)."""
    _require_cuda(backend, "jit_block_mkperm")
    unique = ctypes.c_uint32(0)
    _check(lib().b200_block_mkperm(_stream(stream), _ptr(values), size, block_size,
                                   bucket_count, _ptr(perm), _ptr(offsets),
                                   ctypes.byref(unique)))
    return unique.value


def block_mkperm_async(values, size, block_size, bucket_count, perm, offsets, stream=None):
    """Enqueue only; after a stream sync offsets[4 * bucket_count] is the unique count."""
    _check(lib().b200_block_mkperm_async(_stream(stream), _ptr(values), size, block_size,
                                         bucket_count, _ptr(perm), _ptr(offsets)))


def call_reduce(ids, size, id_bound, perm, offsets, stream=None):
    """The mkperm step of jit_var_call_reduce (src/call.cpp:1268-1389): callable ids in
    [0, id_bound] -> permutation + (id, start, size, 0) records ordered by size (largest
    first, ties by ascending id) in the host-accessible `offsets` (4 * (id_bound + 1) + 1
    words).  Returns the number of records."""
    u = ctypes.c_uint32(0)
    _check(lib().b200_call_reduce(_stream(stream), _ptr(ids), size, id_bound, _ptr(perm),
                                  _ptr(offsets), ctypes.byref(u)))
    return u.value


def mkperm_histogram(values, size, bucket_count, hist, stream=None):
    _check(lib().b200_mkperm_histogram(_stream(stream), _ptr(values), size,
                                       bucket_count, _ptr(hist)))


def scatter_reduce(vt, op, target, value, index, mask, n, mode=ReduceMode.Auto,
                   stream=None):
    """target[index[i]] op= value[i] for i < n where mask[i] (mask may be None)."""
    _check(lib().b200_scatter_reduce(_stream(stream), vt, op, mode, _ptr(target),
                                     _ptr(value), _ptr(index), _ptr(mask), n))


def scatter_inc(target, index, mask, out, n, stream=None):
    """out[i] = target[index[i]]++ (atomically) for i < n where mask[i]; masked
    entries receive 0 (jit_var_scatter_inc, jit.h:1125-1143)."""
    _check(lib().b200_scatter_inc(_stream(stream), _ptr(target), _ptr(index), _ptr(mask),
                                  _ptr(out), n))


def scatter_reduce_packet(vt, op, target, values, index, mask, n, mode=ReduceMode.Auto,
                          stream=None):
    """target[index[i] * W + k] op= values[k][i] for the W = len(values) component
    arrays (jit_var_scatter_packet with a reduction, jit.h:1117)."""
    ptrs = (ctypes.c_void_p * len(values))(*[_ptr(v) for v in values])
    _check(lib().b200_scatter_reduce_packet(_stream(stream), vt, op, mode, _ptr(target), ptrs,
                                            len(values), _ptr(index), _ptr(mask), n))


def scatter_reduce_idx(vt, op, target, value, index, index_vt, mask, n, mode=ReduceMode.Auto,
                       stream=None):
    """scatter_reduce with an index array of type index_vt (Int32 / UInt32 / Int64 /
    UInt64) and with op == Identity (plain scatter) -- jitc_var_scatter, src/op.cpp:2899-3086."""
    _check(lib().b200_scatter_reduce_idx(_stream(stream), vt, op, mode, _ptr(target), _ptr(value),
                                         _ptr(index), index_vt, _ptr(mask), n))


def scatter_packet(vt, target, values, index, mask, n, stream=None):
    """target[index[i] * W + k] = values[k][i] (jit_var_scatter_packet without a
    reduction, src/cuda_packet.cpp:329-443)."""
    ptrs = (ctypes.c_void_p * len(values))(*[_ptr(v) for v in values])
    _check(lib().b200_scatter_packet(_stream(stream), vt, _ptr(target), ptrs, len(values),
                                     _ptr(index), _ptr(mask), n))


def gather_packet(vt, source, outs, index, mask, n, stream=None):
    """outs[k][i] = mask[i] ? source[index[i] * W + k] : 0 (jit_var_gather_packet,
    src/cuda_packet.cpp:18-166)."""
    ptrs = (ctypes.c_void_p * len(outs))(*[_ptr(v) for v in outs])
    _check(lib().b200_gather_packet(_stream(stream), vt, _ptr(source), ptrs, len(outs),
                                    _ptr(index), _ptr(mask), n))


def jit_can_scatter_reduce(backend, vt, op):
    _require_cuda(backend, "jit_can_scatter_reduce")
    return bool(lib().b200_can_scatter_reduce(vt, op))


def jit_all(backend, values, size, stream=None):
    _require_cuda(backend, "jit_all")
    r = ctypes.c_int(0)
    _check(lib().b200_all(_stream(stream), _ptr(values), size, ctypes.byref(r)))
    return bool(r.value)


def jit_any(backend, values, size, stream=None):
    _require_cuda(backend, "jit_any")
    r = ctypes.c_int(0)
    _check(lib().b200_any(_stream(stream), _ptr(values), size, ctypes.byref(r)))
    return bool(r.value)


# ---- multi-GPU: peer-mapped mailboxes (include/drjit_b200.h, "multi-GPU") ------------
class ShardedContext:
    """b200_sharded_*: one context per rank.  `gather_handles(my_handle: bytes) ->
    list[bytes]` exchanges the CUDA IPC handles of the mailboxes (all ranks, rank
    order) over whatever channel the host program has (torch.distributed, MPI, ...)."""

    def __init__(self, rank, world, gather_handles):
        self.rank, self.world = rank, world
        ctx = ctypes.c_void_p()
        _check(lib().b200_sharded_create(rank, world, ctypes.byref(ctx)))
        self._ctx = ctx
        nbytes = lib().b200_sharded_handle_bytes()
        mine = (ctypes.c_uint8 * nbytes)()
        _check(lib().b200_sharded_export(self._ctx, mine))
        handles = gather_handles(bytes(mine))
        assert len(handles) == world and all(len(h) == nbytes for h in handles)
        blob = (ctypes.c_uint8 * (nbytes * world)).from_buffer_copy(b"".join(handles))
        _check(lib().b200_sharded_connect(self._ctx, blob))

    def close(self):
        if self._ctx:
            lib().b200_sharded_destroy(self._ctx)
            self._ctx = None

    def reduce(self, vt, op, in_, local_size, out, stream=None):
        _check(lib().b200_sharded_reduce(self._ctx, _stream(stream), vt, op, _ptr(in_), local_size,
                                         _ptr(out)))

    def reduce_dot(self, vt, a, b, local_size, out, stream=None):
        _check(lib().b200_sharded_reduce_dot(self._ctx, _stream(stream), vt, _ptr(a), _ptr(b), local_size,
                                             _ptr(out)))

    def prefix_reduce(self, vt, op, local_size, exclusive, reverse, in_, out, stream=None):
        _check(lib().b200_sharded_prefix_reduce(self._ctx, _stream(stream), vt, op, local_size,
                                                int(bool(exclusive)), int(bool(reverse)), _ptr(in_),
                                                _ptr(out)))

    def prefix_reduce_cyclic(self, vt, op, local_size, block_size, exclusive, in_, out, stream=None):
        _check(lib().b200_sharded_prefix_reduce_cyclic(self._ctx, _stream(stream), vt, op, local_size,
                                                       block_size, int(bool(exclusive)), _ptr(in_), _ptr(out)))

    def histogram(self, values, local_size, bucket_count, hist, before=None, stream=None):
        _check(lib().b200_sharded_histogram(self._ctx, _stream(stream), _ptr(values), local_size,
                                            bucket_count, _ptr(hist), _ptr(before)))
