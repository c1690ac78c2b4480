"""GPU parity: block reductions / whole-array reductions / dot / all-any through
the C-ABI, against the CPU oracle on the same inputs.  Integer results are
bit-exact; floating point within the tolerance stated at each test."""
import numpy as np
import pytest

import oracle
from cases import RED_SIZES, f32_input, fmix32, red_pairs, u32_input, u64_input
from util import empty_dev, rel_err, to_dev, to_host

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
CUDA = 1


def run_reduce(dr, vt, op, x, bs, offset=0):
    blocks = (x.size + bs - 1) // bs
    d_in = to_dev(x, offset)
    d_out = empty_dev(blocks, x.dtype)
    dr.jit_block_reduce(CUDA, vt, op, x.size, bs, d_in, d_out)
    return to_host(d_out, x.dtype)


@pytest.mark.parametrize("tname", ["u32", "u64"])
def test_block_reduce_grid(dr, O, tname):
    # tests/reductions.cpp:109-151 on the reference's size grid
    bad = []
    for size, bs in red_pairs():
        if size == RED_SIZES[-1] and bs not in (1, 2, 3, 7, 32, 333, 1024, 16384, 169541, size):
            continue
        x = u32_input(size) if tname == "u32" else u64_input(size)
        got = run_reduce(dr, VT[tname], OP["add"], x, bs)
        if not np.array_equal(got, O.block_reduce(VT[tname], OP["add"], x, bs)):
            bad.append((size, bs))
    assert not bad, bad


def test_block_reduce_const(dr):
    # tests/reductions.cpp:109-121 (02_block_reduce_u32_const)
    bad = []
    for size, bs in red_pairs(max_size=170000):
        got = run_reduce(dr, VT["u32"], OP["add"], np.ones(size, dtype=np.uint32), bs)
        blocks = (size + bs - 1) // bs
        expect = np.minimum(size - np.arange(blocks, dtype=np.int64) * bs, bs).astype(np.uint32)
        if not np.array_equal(got, expect):
            bad.append((size, bs))
    assert not bad, bad


INT_TYPES = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}


def int_input(tname, size):
    h = u32_input(size)
    if tname == "u32":
        return h
    if tname == "i32":
        return h.view(np.int32)
    w = (h.astype(np.uint64) << np.uint64(29)) ^ fmix32(h).astype(np.uint64)
    return w if tname == "u64" else w.view(np.int64)


@pytest.mark.parametrize("tname", list(INT_TYPES))
def test_block_reduce_all_int_ops(dr, O, tname):
    bad = []
    sizes = [(1, 1), (5, 2), (37, 37), (1000, 7), (1000, 8), (4096, 64), (4099, 128), (70001, 333),
             (70001, 1024), (70001, 4096), (70001, 20000), (70001, 70001), (300000, 2),
             (300000, 16), (300000, 512), (2000003, 2000003), (2000003, 500)]
    for size, bs in sizes:
        x = int_input(tname, size)
        for opn in ("add", "mul", "min", "max", "and_", "or_"):
            got = run_reduce(dr, VT[tname], OP[opn], x, bs)
            if not np.array_equal(got, O.block_reduce(VT[tname], OP[opn], x, bs)):
                bad.append((size, bs, opn))
    assert not bad, bad


def test_block_reduce_pow2_blocks(dr, O):
    # every power-of-two block size of the benchmark config (1..4096) and beyond
    bad = []
    for size in (1 << 16, (1 << 16) + 37, 1000001):
        x = u32_input(size)
        for lg in range(0, 17):
            bs = 1 << lg
            if bs > size:
                continue
            got = run_reduce(dr, VT["u32"], OP["add"], x, bs)
            if not np.array_equal(got, O.block_reduce(VT["u32"], OP["add"], x, bs)):
                bad.append((size, bs))
            x64 = x.astype(np.uint64)
            got = run_reduce(dr, VT["u64"], OP["max"], x64, bs)
            if not np.array_equal(got, O.block_reduce(VT["u64"], OP["max"], x64, bs)):
                bad.append((size, bs, "u64"))
    assert not bad, bad


def test_block_reduce_misaligned(dr, O):
    bad = []
    for off in (1, 2, 3):
        for size, bs in ((1000, 4), (1000, 7), (5000, 256), (100003, 100003), (100003, 4096),
                         (100003, 33)):
            x = u32_input(size)
            got = run_reduce(dr, VT["u32"], OP["add"], x, bs, offset=off)
            if not np.array_equal(got, O.block_reduce(VT["u32"], OP["add"], x, bs)):
                bad.append((off, size, bs))
    assert not bad, bad


def test_block_reduce_u8(dr, O):
    bad = []
    for size in (4, 5, 64, 1000, 4099, 100001):
        for fill in (0, 1):
            x = np.full(size, fill, dtype=np.uint8)
            x[size // 3] ^= 1
            for opn in ("and_", "or_"):
                for bs in (4, size) if size >= 4 else (size,):
                    got = run_reduce(dr, VT["u8"], OP[opn], x, bs)
                    if not np.array_equal(got, O.block_reduce(VT["u8"], OP[opn], x, bs)):
                        bad.append((size, fill, opn, bs))
    assert not bad, bad


# Floating point: relative error <= 1e-5 against the fp64-accumulated oracle
# (SURVEY.md section 8d); inputs uniform in [0, 1).
@pytest.mark.parametrize("tname,tol", [("f32", 1e-5), ("f64", 1e-12)])
def test_block_reduce_float_add(dr, O, tname, tol):
    bad = []
    for size, bs in ((1000, 7), (4096, 4), (65536, 256), (100003, 1024), (100003, 100003),
                     (1 << 20, 1 << 20), (1 << 20, 4096), (1 << 20, 2), (3000017, 3000017)):
        x = f32_input(size).astype(oracle.NP_OF_VT[VT[tname]])
        got = run_reduce(dr, VT[tname], OP["add"], x, bs)
        ref = O.block_reduce(VT[tname], OP["add"], x, bs, wide=True)
        err = rel_err(got, ref)
        if err > tol:
            bad.append((size, bs, err))
    assert not bad, bad


def test_block_reduce_float_minmax_mul(dr, O):
    bad = []
    for tname in ("f32", "f64", "f16"):
        dt = oracle.NP_OF_VT[VT[tname]]
        for size, bs in ((1000, 7), (100003, 1024), (100003, 100003), (65536, 16)):
            x = (f32_input(size) * 4 - 2).astype(dt)
            for opn in ("min", "max"):  # exact
                got = run_reduce(dr, VT[tname], OP[opn], x, bs)
                if not np.array_equal(got, O.block_reduce(VT[tname], OP[opn], x, bs)):
                    bad.append((tname, size, bs, opn))
        # products of short blocks (values near 1 so nothing over/underflows)
        x = (1 + (f32_input(4096) - 0.5) * 0.01).astype(dt)
        got = run_reduce(dr, VT[tname], OP["mul"], x, 8)
        ref = O.block_reduce(VT[tname], OP["mul"], x, 8, wide=True)
        if rel_err(got, ref) > (2e-3 if tname == "f16" else 1e-5):
            bad.append((tname, "mul", rel_err(got, ref)))
    assert not bad, bad


def test_block_reduce_f16_add(dr, O):
    # f16 accumulates in f32 and narrows once: <= 1 ulp(f16) = 2^-10 relative
    bad = []
    for size, bs in ((1000, 8), (4096, 64), (100003, 1024), (50000, 50000)):
        x = (f32_input(size) * 0.125).astype(np.float16)
        got = run_reduce(dr, VT["f16"], OP["add"], x, bs)
        ref = O.block_reduce(VT["f16"], OP["add"], x, bs, wide=True)
        err = rel_err(got.astype(np.float64), ref.astype(np.float64))
        if err > 2.0 ** -10:
            bad.append((size, bs, err))
    assert not bad, bad


def test_reduce_entry_point_and_determinism(dr, O):
    x = f32_input(1 << 22)
    d_in = to_dev(x)
    d_out = empty_dev(1, np.float32)
    dr.jit_reduce(CUDA, VT["f32"], OP["add"], d_in, x.size, d_out)
    first = to_host(d_out, np.float32).copy()
    for _ in range(3):  # fixed combination order -> bitwise reproducible
        dr.jit_reduce(CUDA, VT["f32"], OP["add"], d_in, x.size, d_out)
        assert to_host(d_out, np.float32)[0] == first[0]
    assert rel_err(first, O.block_reduce(VT["f32"], OP["add"], x, x.size, wide=True)) <= 1e-5


def test_reduce_dot(dr, O):
    bad = []
    for tname, tol in (("f32", 1e-5), ("f64", 1e-12), ("f16", 2.0 ** -9)):
        dt = oracle.NP_OF_VT[VT[tname]]
        for size in (1, 7, 1000, 100003, 1 << 21):
            a = (f32_input(size) * (0.05 if tname == "f16" else 1)).astype(dt)
            b = (f32_input(size, salt=77) * (0.05 if tname == "f16" else 1)).astype(dt)
            d_out = empty_dev(1, dt)
            dr.jit_reduce_dot(CUDA, VT[tname], to_dev(a), to_dev(b), size, d_out)
            got = to_host(d_out, dt).astype(np.float64)
            ref = np.dot(a.astype(np.float64), b.astype(np.float64))
            if abs(got[0] - ref) > tol * max(abs(ref), 1e-30):
                bad.append((tname, size, got[0], ref))
    assert not bad, bad


def test_all_any(dr, O):
    # tests/reductions.cpp:78-107 (01_all_any)
    bad = []
    for i in (0, 1, 2, 3, 5, 11, 20, 37):
        size = 23 * i * i * i + 1
        f = np.zeros(size, dtype=np.uint8)
        t = np.ones(size, dtype=np.uint8)
        for arr in (f, t):
            d = to_dev(arr, offset_elems=i % 4)
            if dr.jit_all(CUDA, d, size) != O.all(arr) or dr.jit_any(CUDA, d, size) != O.any(arr):
                bad.append((size, "const", int(arr[0])))
        for pos in (0, size // 2, size - 1):
            f2, t2 = f.copy(), t.copy()
            f2[pos] = 1
            t2[pos] = 0
            for arr in (f2, t2):
                d = to_dev(arr)
                if dr.jit_all(CUDA, d, size) != O.all(arr) or dr.jit_any(CUDA, d, size) != O.any(arr):
                    bad.append((size, pos))
    assert not bad, bad


def test_errors(dr):
    # src/cuda_ts.cpp:201-207, :313-315
    d = to_dev(np.ones(16, dtype=np.uint32))
    o = empty_dev(16, np.uint32)
    with pytest.raises(RuntimeError, match="invalid block size"):
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], 16, 0, d, o)
    with pytest.raises(RuntimeError, match="invalid block size"):
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], 16, 17, d, o)
    with pytest.raises(RuntimeError, match="no existing kernel"):
        dr.jit_block_reduce(CUDA, VT["f32"], OP["and_"], 16, 4, d, o)
    with pytest.raises(RuntimeError, match="no existing kernel"):
        dr.jit_block_reduce(CUDA, VT["u16"], OP["add"], 16, 4, d, o)
    with pytest.raises(RuntimeError):
        dr.jit_block_reduce(2, VT["u32"], OP["add"], 16, 4, d, o)  # LLVM backend: not provided
    dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], 0, 0, d, o)  # size == 0: silent no-op


def test_full_size_reduce(dr, O):
    # BASELINE.json configs[0]/[1] size: 2^28 elements, bit-exact for u32
    n = 1 << 28
    x = u32_input(n)
    d_in = to_dev(x)
    d_out = empty_dev(1, np.uint32)
    dr.jit_reduce(CUDA, VT["u32"], OP["add"], d_in, n, d_out)
    assert to_host(d_out, np.uint32)[0] == O.block_reduce(VT["u32"], OP["add"], x, n)[0]
    for bs in (2, 4096, 1 << 20):
        o = empty_dev(n // bs, np.uint32)
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], n, bs, d_in, o)
        assert np.array_equal(to_host(o, np.uint32), O.block_reduce(VT["u32"], OP["add"], x, bs))
    # fp32 view of the same bits is not meaningful; use the C2 generator
    del d_in
    xf = f32_input(n)
    d_f = to_dev(xf)
    dr.jit_reduce(CUDA, VT["f32"], OP["add"], d_f, n, d_out)
    got = to_host(d_out, np.float32)[0]
    assert abs(got - xf.astype(np.float64).sum()) <= 1e-5 * xf.astype(np.float64).sum()
