"""GPU parity at the BASELINE.json sizes (2^28 reduce / scan / compress, 2^26
mkperm, 2^26 -> 2^20 scatter) against the REFERENCE'S OWN CUDA KERNELS
(oracle/_ref/libref_cuda.so on the same device buffers), plus the per-prefix
fp32 check of the whole-array scan that the bench times.

Everything is compared on the device (the arrays are 1 GiB each); inputs are
the counter-based generators of SURVEY.md section 8d (C1-C5), produced on the
device with the same fmix32 as tests/golden/cases.py.

Tolerances: integers bit-exact; fp32 Add <= 1e-5 relative to an fp64
accumulation and <= 2e-5 against the reference CUDA result; mkperm: unique
count, offsets records as a set, per-bucket index sets.
"""
import numpy as np
import pytest
import torch

import oracle
from util import dev_f32_input, dev_fmix32, dev_u32_input

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
CUDA = 1
N28 = 1 << 28
N26 = 1 << 26


@pytest.fixture(scope="module")
def R(dr):
    if not oracle.ref_cuda_available():
        pytest.skip("oracle/_ref/libref_cuda.so has not been built")
    return oracle.ReferenceCUDA.get()


def ref_call(R, fn, *args):
    torch.cuda.synchronize()
    r = fn(*args)
    R.sync()
    return r


def test_scan_f32_2_28_every_prefix(dr):
    # the call `scan_bsN` of bench.py: whole-array exclusive fp32 prefix sum over
    # 2^28 elements (chained streaming kernel), EVERY prefix within 1e-5 of an fp64
    # accumulation on the host
    x = dev_f32_input(N28)
    out = torch.empty_like(x)
    dr.jit_block_prefix_reduce(CUDA, VT["f32"], OP["add"], N28, N28, 1, 0, x, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = np.cumsum(x.cpu().numpy(), dtype=np.float64)
    assert got[0] == 0.0
    err = np.abs(got[1:].astype(np.float64) - ref[:-1]) / np.maximum(ref[:-1], 1.0)
    assert float(err.max()) <= 1e-5, float(err.max())
    del got, ref, err
    # inclusive + reverse through the same kernel
    dr.jit_block_prefix_reduce(CUDA, VT["f32"], OP["add"], N28, N28, 0, 1, x, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = np.cumsum(x.cpu().numpy()[::-1], dtype=np.float64)[::-1]
    err = np.abs(got.astype(np.float64) - ref) / np.maximum(ref, 1.0)
    assert float(err.max()) <= 1e-5, float(err.max())


def test_reduce_scan_u32_2_28_vs_reference_cuda(dr, R):
    x = dev_u32_input(N28)
    a, b = torch.empty_like(x), torch.empty_like(x)
    for bs in (N28, 4096, 1000):
        nb = -(-N28 // bs)
        dr.jit_block_reduce(CUDA, VT["u32"], OP["add"], N28, bs, x, a)
        ref_call(R, R.block_reduce, VT["u32"], OP["add"], N28, bs, x.data_ptr(), b.data_ptr())
        assert torch.equal(a[:nb], b[:nb]), ("reduce", bs)
    for bs, excl, rev in ((N28, 1, 0), (N28, 0, 1), (4096, 1, 0), (1 << 20, 0, 0)):
        dr.jit_block_prefix_reduce(CUDA, VT["u32"], OP["add"], N28, bs, excl, rev, x, a)
        ref_call(R, R.block_prefix_reduce, VT["u32"], OP["add"], N28, bs, excl, rev, x.data_ptr(),
                 b.data_ptr())
        assert torch.equal(a, b), ("scan", bs, excl, rev)


def test_reduce_scan_f32_2_28_vs_reference_cuda(dr, R):
    x = dev_f32_input(N28)
    a, b = torch.empty_like(x), torch.empty_like(x)
    for bs in (N28, 4096, 2):
        nb = -(-N28 // bs)
        dr.jit_block_reduce(CUDA, VT["f32"], OP["add"], N28, bs, x, a)
        ref_call(R, R.block_reduce, VT["f32"], OP["add"], N28, bs, x.data_ptr(), b.data_ptr())
        err = ((a[:nb].double() - b[:nb].double()).abs() / b[:nb].double().abs().clamp_min(1e-30)).max()
        assert float(err) <= 2e-5, ("reduce", bs, float(err))
    for bs, excl in ((N28, 1), (4096, 1), (64, 0)):
        dr.jit_block_prefix_reduce(CUDA, VT["f32"], OP["add"], N28, bs, excl, 0, x, a)
        ref_call(R, R.block_prefix_reduce, VT["f32"], OP["add"], N28, bs, excl, 0, x.data_ptr(),
                 b.data_ptr())
        # relative to max(|ref|, 1): an exclusive prefix starts at 0, inputs are in [0, 1)
        err = ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1.0)).max()
        assert float(err) <= 2e-5, ("scan", bs, excl, float(err))


@pytest.mark.parametrize("density", [0.01, 0.5, 0.99])
def test_compress_2_28_vs_reference_cuda(dr, R, density):
    h = dev_fmix32(N28, xor=0x9E3779B9)
    thr = min(int(density * 2 ** 32), 2 ** 32)
    mask = torch.zeros(2 * N28, dtype=torch.uint8, device="cuda")  # the reference zero-fills past size
    chunk = 1 << 26
    for s in range(0, N28, chunk):
        mask[s:s + chunk] = ((h[s:s + chunk].to(torch.int64) & 0xFFFFFFFF) < thr).to(torch.uint8)
    del h
    a = torch.empty(N28, dtype=torch.int32, device="cuda")
    b = torch.empty(N28, dtype=torch.int32, device="cuda")
    ca = dr.jit_compress(CUDA, mask, N28, a)
    cb = ref_call(R, R.compress, mask.data_ptr(), N28, b.data_ptr())
    assert ca == cb
    assert torch.equal(a[:ca], b[:cb])


def _bucket_canonical(keys, perm):
    """(bucket, index) pairs along a permutation, sorted -> identical for two
    permutations iff every bucket holds the same index SET"""
    p = perm.to(torch.int64) & 0xFFFFFFFF
    k = keys[p].to(torch.int64) & 0xFFFFFFFF
    assert bool((k[1:] >= k[:-1]).all()), "keys are not grouped in ascending bucket order"
    return torch.sort((k << 32) | p).values


@pytest.mark.parametrize("buckets", [16, 1024, 65536])
def test_mkperm_2_26_vs_reference_cuda(dr, R, buckets):
    for skew in (False, True):
        h = dev_fmix32(N26).to(torch.int64) & 0xFFFFFFFF
        keys = h % buckets
        if skew:  # tests/golden/cases.py key_input(skew=True)
            i = torch.arange(N26, dtype=torch.int64, device="cuda")
            h2 = dev_fmix32(N26, index=i >> 1).to(torch.int64) & 0xFFFFFFFF
            keys = torch.where((h % 100) < 90, torch.full_like(h, min(1, buckets - 1)), h2 % buckets)
        keys = keys.to(torch.int32)
        del h
        a = torch.empty(N26, dtype=torch.int32, device="cuda")
        b = torch.empty(N26, dtype=torch.int32, device="cuda")
        oa = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
        ob = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
        ua = dr.jit_block_mkperm(CUDA, keys, N26, N26, buckets, a, oa)
        ub = ref_call(R, R.block_mkperm, keys.data_ptr(), N26, N26, buckets, b.data_ptr(), ob.data_ptr())
        assert ua == ub, (skew, ua, ub)
        ra = {tuple(r) for r in oa.numpy().view(np.uint32)[:4 * ua].reshape(ua, 4).tolist()}
        rb = {tuple(r) for r in ob.numpy().view(np.uint32)[:4 * ub].reshape(ub, 4).tolist()}
        assert ra == rb, (skew, "records")
        ca = _bucket_canonical(keys, a)
        cb = _bucket_canonical(keys, b)
        assert torch.equal(ca, cb), (skew, "index sets")
        # ours is additionally stable: the canonical order IS the permutation
        assert torch.equal(ca & 0xFFFFFFFF, a.to(torch.int64) & 0xFFFFFFFF), (skew, "stability")
        del ca, cb


@pytest.mark.parametrize("kind", ["random", "coherent"])
def test_scatter_add_2_26_to_2_20_vs_reference_cuda(dr, R, kind):
    m = 1 << 20
    if kind == "random":
        idx = ((dev_fmix32(N26).to(torch.int64) & 0xFFFFFFFF) % m).to(torch.int32)
    else:
        idx = (torch.arange(N26, dtype=torch.int64, device="cuda") >> 6).to(torch.int32)
    # u32: bit-exact
    val = dev_u32_input(N26)
    for mode in (0, 1, 2):
        a = torch.zeros(m, dtype=torch.int32, device="cuda")
        b = torch.zeros(m, dtype=torch.int32, device="cuda")
        dr.scatter_reduce(VT["u32"], OP["add"], a, val, idx, None, N26, mode=mode)
        ref_call(R, R.scatter_reduce, VT["u32"], OP["add"], mode, b.data_ptr(), m, val.data_ptr(),
                 idx.data_ptr(), None, N26)
        assert torch.equal(a, b), ("u32", mode)
    # f32 Add: <= 2e-5 against the reference CUDA result, <= 1e-5 against fp64
    valf = dev_f32_input(N26)
    ref64 = torch.zeros(m, dtype=torch.float64, device="cuda")
    ref64.index_add_(0, idx.to(torch.int64), valf.double())
    for mode in (0, 1, 2):
        a = torch.zeros(m, dtype=torch.float32, device="cuda")
        b = torch.zeros(m, dtype=torch.float32, device="cuda")
        dr.scatter_reduce(VT["f32"], OP["add"], a, valf, idx, None, N26, mode=mode)
        ref_call(R, R.scatter_reduce, VT["f32"], OP["add"], mode, b.data_ptr(), m, valf.data_ptr(),
                 idx.data_ptr(), None, N26)
        e_ref = ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1e-30)).max()
        e_64 = ((a.double() - ref64).abs() / ref64.abs().clamp_min(1e-30)).max()
        assert float(e_ref) <= 2e-5 and float(e_64) <= 1e-5, ("f32", mode, float(e_ref), float(e_64))
