"""GPU: the JitFlag bits and interop calls of the boundary (SURVEY.md section 8b
"Ordering / sync", section 5) through the C-ABI: KernelHistory, LaunchBlocking,
ForbidSynchronization (jit.h:1734-1742), jit_cuda_sync_stream (jit.h:243-255),
jit_malloc_migrate (jit.h:516), the stream-ordered release of jit_free."""
import numpy as np
import pytest
import torch

import oracle
from cases import mask_input, u32_input
from util import empty_dev, to_dev, to_host

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
CUDA = 1


def test_kernel_history(dr):
    n = 1 << 20
    x = to_dev(u32_input(n))
    out = empty_dev(n, np.uint32)
    mask = to_dev(mask_input(n, 0.5))
    dr.jit_kernel_history_clear()
    dr.jit_set_flag(dr.JitFlag.KernelHistory, True)
    try:
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, 1024, x, out)
        dr.jit_block_prefix_reduce(CUDA, VT["u32"], OP["add"], n, n, 1, 0, x, out)
        dr.jit_reduce_dot(CUDA, VT["f32"], x, x, n, out)
        dr.jit_compress(CUDA, mask, n, out)
        dr.jit_block_mkperm(CUDA, x, n, n, 0xFFFFFFFF, out, None)
        dr.jit_memset_async(CUDA, out, n, 4, (0x01020304).to_bytes(4, "little"))
    finally:
        dr.jit_set_flag(dr.JitFlag.KernelHistory, False)
    hist = dr.jit_kernel_history()
    types = [h[0] for h in hist]
    K = dr.KernelType
    for t in (K.BlockReduce, K.BlockPrefixReduce, K.Dot, K.Compress, K.MkPerm, K.Memset):
        assert t in types, (t, types)
    # (primitives that call other primitives -- mkperm scans its count table -- add
    # entries of their own, as in the reference)
    assert all(h[2] >= 0.0 for h in hist)
    for t in (K.BlockReduce, K.Dot, K.Compress, K.MkPerm, K.Memset):
        assert any(h[0] == t and h[1] == n for h in hist), t
    assert dr.jit_kernel_history() == []  # reading clears
    dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, 1024, x, out)
    assert dr.jit_kernel_history() == []  # flag off: nothing recorded


def test_launch_blocking(dr):
    n = 1 << 24
    x = to_dev(u32_input(n))
    pinned = torch.zeros(16, dtype=torch.int32).pin_memory()
    expect = int(u32_input(n).sum(dtype=np.uint32))
    dr.jit_set_flag(dr.JitFlag.LaunchBlocking, True)
    try:
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, n, x, pinned.data_ptr())
        got = int(pinned.numpy().view(np.uint32)[0])  # no synchronisation on purpose
    finally:
        dr.jit_set_flag(dr.JitFlag.LaunchBlocking, False)
    assert got == expect


def test_forbid_synchronization(dr):
    n = 100000
    x = to_dev(u32_input(n))
    out = empty_dev(n, np.uint32)
    mask = to_dev(mask_input(n, 0.5))
    dr.jit_set_flag(dr.JitFlag.ForbidSynchronization, True)
    try:
        for call in (lambda: dr.jit_compress(CUDA, mask, n, out),
                     lambda: dr.jit_sync_thread(),
                     lambda: dr.jit_all(CUDA, mask, n),
                     lambda: dr.jit_memcpy(CUDA, out, x, 4 * n)):
            with pytest.raises(RuntimeError, match="forbidden"):
                call()
        # asynchronous entry points keep working
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, n, x, out)
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        dr.compress_async(mask, n, out, cnt)
    finally:
        dr.jit_set_flag(dr.JitFlag.ForbidSynchronization, False)
    assert int(cnt.item()) == int(mask_input(n, 0.5).sum())
    assert dr.jit_flags() & (dr.JitFlag.ForbidSynchronization | dr.JitFlag.LaunchBlocking) == 0


def test_library_stream_orders_with_the_null_stream(dr):
    # the library stream is a BLOCKING stream like the reference's (src/cuda_core.cpp:480):
    # inputs staged on the legacy default stream are seen, results are readable from it
    n = 1 << 24
    lib_stream = dr.jit_cuda_stream()
    assert lib_stream
    for _ in range(5):
        x = torch.randint(0, 1 << 20, (n,), dtype=torch.int32, device="cuda")  # stream 0
        out = torch.zeros(4, dtype=torch.int32, device="cuda")
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, n, x, out, stream=lib_stream)
        got = int(out[0].item()) & 0xFFFFFFFF  # stream 0 again, no explicit synchronisation
        assert got == int(x.sum(dtype=torch.int64).item()) & 0xFFFFFFFF


def test_cuda_sync_stream_and_migrate(dr):
    n = 1 << 22
    h = u32_input(n)
    d = dr.jit_malloc(CUDA, 4 * n)
    dr.jit_memcpy(CUDA, d, h, 4 * n)
    out = dr.jit_malloc(CUDA, 64)
    lib_stream = dr.jit_cuda_stream()
    dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, n, d, out, stream=lib_stream)
    side = torch.cuda.Stream()  # non-blocking stream: needs the explicit hand-over
    dr.jit_cuda_sync_stream(side.cuda_stream)
    res = torch.zeros(1, dtype=torch.int32).pin_memory()
    dr.jit_memcpy_async(CUDA, res.data_ptr(), out, 4, stream=side)
    side.synchronize()
    assert int(res.numpy().view(np.uint32)[0]) == int(h.sum(dtype=np.uint32))
    # device -> pinned host copy, then moved back
    host = dr.jit_malloc_migrate(d, dr.JitBackend.None_, move=0)
    dr.jit_sync_thread(stream=lib_stream)
    import ctypes
    got = np.ctypeslib.as_array((ctypes.c_uint32 * n).from_address(host))
    assert np.array_equal(got, h)
    back = dr.jit_malloc_migrate(host, CUDA, move=1)
    dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, n, back, out, stream=lib_stream)
    dr.jit_sync_thread(stream=lib_stream)
    t = torch.zeros(1, dtype=torch.int32, device="cuda")
    dr.jit_memcpy(CUDA, t, out, 4)
    assert int(t.item()) & 0xFFFFFFFF == int(h.sum(dtype=np.uint32))
    for p in (d, back, out):
        dr.jit_free(p, stream=lib_stream)
