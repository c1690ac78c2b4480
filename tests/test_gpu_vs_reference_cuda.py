"""GPU parity against the REFERENCE'S OWN CUDA KERNELS: the unmodified reference
built with its CUDA backend (oracle/_ref/libref_cuda.so, compute_75 PTX JIT
compiled by the driver for the B200) runs on the same device buffers as the
sm_100a kernels of this repository.

Bar (BASELINE.json north_star / SURVEY.md section 8d):
  * integer reduce / scan / compress / scatter: bit-exact;
  * mkperm: unique count, offsets records as a SET, per-bucket index SETS (the
    reference's CUDA permutation is not stable and its record order is
    nondeterministic, mkperm.cuh:302-317);
  * fp32 Add reduce / scan / scatter: relative error <= 2e-5 against the
    reference CUDA result (all-positive inputs); fp Min/Max exact.
"""
import numpy as np
import pytest
import torch

import oracle
from cases import (RED_SIZES, cubic_sizes, f32_input, fmix32, key_input, mask_input, u32_input,
                   u64_input)
from util import empty_dev, rel_err, to_dev, to_host

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
CUDA = 1
FP32_TOL = 2e-5


@pytest.fixture(scope="module")
def R(dr):
    if not oracle.ref_cuda_available():
        pytest.skip("oracle/_ref/libref_cuda.so has not been built")
    return oracle.ReferenceCUDA.get()


def ref_call(R, fn, *args):
    """Run a reference call after torch's work is done and wait for it."""
    torch.cuda.synchronize()
    r = fn(*args)
    R.sync()
    return r


INPUTS = {"u32": u32_input, "u64": u64_input,
          "i32": lambda n: u32_input(n).view(np.int32),
          "i64": lambda n: (u64_input(n) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64),
          "f32": f32_input}


@pytest.mark.parametrize("tname,opn", [("u32", "add"), ("u64", "add"), ("i32", "min"), ("i32", "max"),
                                       ("u32", "and_"), ("u64", "or_"), ("i64", "max"),
                                       ("u32", "mul"), ("f32", "add"), ("f32", "min"), ("f32", "max")])
def test_block_reduce_vs_reference_cuda(dr, R, tname, opn):
    vt, op = VT[tname], OP[opn]
    dt = oracle.NP_OF_VT[vt]
    bad = []
    for size in RED_SIZES + [(1 << 22) + 7]:
        x = INPUTS[tname](size)
        d_x = to_dev(x)
        for bs in [b for b in RED_SIZES if b <= size] + [size]:
            nb = (size + bs - 1) // bs
            d_a, d_b = empty_dev(nb, dt), empty_dev(nb, dt)
            dr.jit_block_reduce(CUDA, vt, op, size, bs, d_x, d_a)
            ref_call(R, R.block_reduce, vt, op, size, bs, d_x.data_ptr(), d_b.data_ptr())
            a, b = to_host(d_a, dt), to_host(d_b, dt)
            if tname == "f32" and opn == "add":
                if rel_err(a, b) > FP32_TOL:
                    bad.append((size, bs, rel_err(a, b)))
            elif not np.array_equal(a, b):
                bad.append((size, bs))
    assert not bad, bad[:10]


@pytest.mark.parametrize("tname,opn", [("u32", "add"), ("u64", "add"), ("i32", "min"), ("u32", "or_"),
                                       ("f32", "add"), ("f32", "max")])
def test_block_prefix_reduce_vs_reference_cuda(dr, R, tname, opn):
    vt, op = VT[tname], OP[opn]
    dt = oracle.NP_OF_VT[vt]
    bad = []
    sizes = RED_SIZES if tname != "f32" else [s for s in RED_SIZES if s in (7, 333, 16384, 9973 * 17)]
    for size in sizes + [(1 << 22) + 7]:
        x = INPUTS[tname](size)
        d_x = to_dev(x)
        d_a, d_b = empty_dev(size, dt), empty_dev(size, dt)
        for bs in [b for b in RED_SIZES if b <= size] + [size]:
            for excl in (0, 1):
                for rev in (0, 1):
                    dr.jit_block_prefix_reduce(CUDA, vt, op, size, bs, excl, rev, d_x, d_a)
                    ref_call(R, R.block_prefix_reduce, vt, op, size, bs, excl, rev,
                             d_x.data_ptr(), d_b.data_ptr())
                    a, b = to_host(d_a, dt), to_host(d_b, dt)
                    if tname == "f32" and opn == "add":
                        # absolute tolerance scaled by the block total: an exclusive
                        # prefix starts at 0 and the inputs are positive
                        scale = max(1.0, float(np.max(np.abs(b.astype(np.float64)))))
                        err = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / scale
                        if err > FP32_TOL:
                            bad.append((size, bs, excl, rev, err))
                    elif not np.array_equal(a, b):
                        bad.append((size, bs, excl, rev))
    assert not bad, bad[:10]


def test_reduce_dot_vs_reference_cuda(dr, R):
    for size in (1, 1000, 16384, (1 << 22) + 7):
        a, b = f32_input(size), f32_input(size, salt=77)
        d_a, d_b = to_dev(a), to_dev(b)
        d_o, d_r = empty_dev(1, np.float32), empty_dev(1, np.float32)
        dr.jit_reduce_dot(CUDA, VT["f32"], d_a, d_b, size, d_o)
        ref_call(R, R.reduce_dot, VT["f32"], d_a.data_ptr(), d_b.data_ptr(), size, d_r.data_ptr())
        assert rel_err(to_host(d_o, np.float32), to_host(d_r, np.float32)) <= FP32_TOL


def test_compress_vs_reference_cuda(dr, R):
    bad = []
    for size in cubic_sizes(30)[1:] + [(1 << 24) + 3]:
        for dens in (0.0, 0.01, 0.5, 0.99, 1.0):
            m = mask_input(size, dens)
            # the reference zero-fills the mask buffer up to a power of two
            buf = torch.zeros(2 * size + 8192, dtype=torch.uint8, device="cuda")
            buf[:size].copy_(torch.from_numpy(m))
            d_a, d_b = empty_dev(size, np.uint32), empty_dev(size, np.uint32)
            ca = dr.jit_compress(CUDA, buf, size, d_a)
            cb = ref_call(R, R.compress, buf.data_ptr(), size, d_b.data_ptr())
            if ca != cb or not np.array_equal(to_host(d_a, np.uint32)[:ca], to_host(d_b, np.uint32)[:cb]):
                bad.append((size, dens, ca, cb))
    assert not bad, bad[:10]


def _bucket_sets(perm, offs, uq):
    """{bucket id: sorted index array} from a permutation and its offsets records"""
    rec = offs[:4 * uq].reshape(uq, 4)
    return {int(r[0]): np.sort(perm[r[1]:r[1] + r[2]]) for r in rec}


@pytest.mark.parametrize("buckets", [1, 2, 16, 100, 1024, 5000, 65536])
def test_mkperm_vs_reference_cuda(dr, R, buckets):
    bad = []
    for size in (1, 1000, 23 * 20 ** 3 + 1, (1 << 22) + 5):
        for skew in (False, True):
            k = key_input(size, buckets, skew)
            d_k = to_dev(k)
            d_a, d_b = empty_dev(size, np.uint32), empty_dev(size, np.uint32)
            oa = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
            ob = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
            ua = dr.jit_block_mkperm(CUDA, d_k, size, size, buckets, d_a, oa)
            ub = ref_call(R, R.block_mkperm, d_k.data_ptr(), size, size, buckets, d_b.data_ptr(),
                          ob.data_ptr())
            if ua != ub:
                bad.append((size, skew, "unique", ua, ub))
                continue
            sa = _bucket_sets(to_host(d_a, np.uint32), oa.numpy().view(np.uint32), ua)
            sb = _bucket_sets(to_host(d_b, np.uint32), ob.numpy().view(np.uint32), ub)
            if sa.keys() != sb.keys() or any(not np.array_equal(sa[b], sb[b]) for b in sa):
                bad.append((size, skew, "sets"))
            # records as a set: (id, start, size) -- starts agree because both
            # lay the buckets out in ascending id order
            ra = {tuple(r) for r in oa.numpy().view(np.uint32)[:4 * ua].reshape(ua, 4).tolist()}
            rb = {tuple(r) for r in ob.numpy().view(np.uint32)[:4 * ub].reshape(ub, 4).tolist()}
            if ra != rb:
                bad.append((size, skew, "records"))
    assert not bad, bad[:10]


def _index(n, m, kind):
    i = np.arange(n, dtype=np.uint32)
    if kind == "random":
        return (fmix32(i) % np.uint32(m)).astype(np.uint32)
    return ((i >> np.uint32(6)) % np.uint32(m)).astype(np.uint32)


@pytest.mark.parametrize("tname,opn", [("u32", "add"), ("i32", "min"), ("u32", "max"), ("u64", "add"),
                                       ("u32", "and_"), ("u32", "or_"), ("f32", "add"), ("f32", "min"),
                                       ("f32", "max")])
def test_scatter_vs_reference_cuda(dr, R, tname, opn):
    vt, op = VT[tname], OP[opn]
    dt = oracle.NP_OF_VT[vt]
    assert R.can_scatter_reduce(vt, op) == dr.jit_can_scatter_reduce(CUDA, vt, op)
    bad = []
    for n, m in ((1000, 7), (100003, 997), (1 << 22, 1 << 16)):
        val = INPUTS[tname](n)
        if tname == "f32" and opn != "add":
            val = (val - np.float32(0.5)).astype(np.float32)  # both signs
        ident = oracle.Oracle().reduce_identity(vt, op)
        if dt().itemsize == 4:
            tgt = np.full(m, ident & 0xFFFFFFFF, dtype=np.uint32).view(dt)
        else:
            tgt = np.full(m, ident, dtype=np.uint64).view(dt)
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        for kind in ("random", "coherent"):
            idx = _index(n, m, kind)
            d_v, d_i, d_m = to_dev(val), to_dev(idx), to_dev(mask)
            for mk in (None, d_m):
                for mode in (0, 1, 2):
                    d_a, d_b = to_dev(tgt), to_dev(tgt)
                    dr.scatter_reduce(vt, op, d_a, d_v, d_i, mk, n, mode=mode)
                    ref_call(R, R.scatter_reduce, vt, op, mode, d_b.data_ptr(), m, d_v.data_ptr(),
                             d_i.data_ptr(), None if mk is None else mk.data_ptr(), n)
                    a, b = to_host(d_a, dt), to_host(d_b, dt)
                    if tname == "f32" and opn == "add":
                        if rel_err(a, b) > FP32_TOL:
                            bad.append((n, m, kind, mode, mk is not None, rel_err(a, b)))
                    elif not np.array_equal(a, b):
                        bad.append((n, m, kind, mode, mk is not None))
    assert not bad, bad[:10]


def test_scatter_inc_vs_reference_cuda(dr, R):
    # jit_var_scatter_inc through the reference's JIT vs b200_scatter_inc: which entry
    # of a counter gets which old value is unspecified on both sides, so both results
    # must satisfy the contract (oracle.scatter_inc_check) and leave the SAME counters
    O = oracle.Oracle()
    bad = []
    for n, m in ((1000, 1), (100003, 7), (1 << 20, 1 << 10)):
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        before = (fmix32(np.arange(m, dtype=np.uint32)) & np.uint32(0xffff)).astype(np.uint32)
        for kind in ("random", "coherent"):
            idx = _index(n, m, kind)
            d_i, d_m = to_dev(idx), to_dev(mask)
            for mk, hmask in ((None, None), (d_m, mask)):
                d_a, d_b = to_dev(before), to_dev(before)
                o_a, o_b = empty_dev(n, np.uint32), empty_dev(n, np.uint32)
                dr.scatter_inc(d_a, d_i, mk, o_a, n)
                ref_call(R, R.scatter_inc, d_b.data_ptr(), m, d_i.data_ptr(),
                         None if mk is None else mk.data_ptr(), n, o_b.data_ptr())
                a, b = to_host(d_a, np.uint32), to_host(d_b, np.uint32)
                va = O.scatter_inc_check(before, a, idx, hmask, to_host(o_a, np.uint32))
                ob = to_host(o_b, np.uint32).copy()
                if hmask is not None:
                    ob[hmask == 0] = 0  # undefined in the reference (jit.h:1136)
                vb = O.scatter_inc_check(before, b, idx, hmask, ob)
                if va or vb or not np.array_equal(a, b):
                    bad.append((n, m, kind, mk is not None, va, vb))
    assert not bad, bad[:10]


@pytest.mark.parametrize("width", [2, 4])
def test_scatter_packet_vs_reference_cuda(dr, R, width):
    # jit_var_scatter_packet (reducing) through the reference's JIT vs
    # b200_scatter_reduce_packet: u32 bit-exact, f32 Add within FP32_TOL, f32 Max exact
    bad = []
    for n, m in ((1000, 7), (100003, 997), (1 << 20, 1 << 14)):
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        for kind in ("random", "coherent"):
            idx = _index(n, m, kind)
            d_i, d_m = to_dev(idx), to_dev(mask)
            for tname, opn in (("u32", "add"), ("u32", "max"), ("f32", "add"), ("f32", "max")):
                vt, op = VT[tname], OP[opn]
                dt = oracle.NP_OF_VT[vt]
                if tname == "u32":
                    vals = [(u32_input(n) >> np.uint32(8 + k)).astype(np.uint32) for k in range(width)]
                else:
                    vals = [(f32_input(n) + np.float32(k) - np.float32(0.5 if opn == "max" else 0.0)).astype(np.float32)
                            for k in range(width)]
                ident = oracle.Oracle().reduce_identity(vt, op)
                tgt = np.full(m * width, ident & 0xFFFFFFFF, dtype=np.uint32).view(dt)
                d_vals = [to_dev(v) for v in vals]
                for mk in (None, d_m):
                    for mode in (0, 1):
                        d_a, d_b = to_dev(tgt), to_dev(tgt)
                        dr.scatter_reduce_packet(vt, op, d_a, d_vals, d_i, mk, n, mode=mode)
                        ref_call(R, R.scatter_packet, vt, op, mode, d_b.data_ptr(), m * width,
                                 [v.data_ptr() for v in d_vals], d_i.data_ptr(),
                                 None if mk is None else mk.data_ptr(), n)
                        a, b = to_host(d_a, dt), to_host(d_b, dt)
                        if tname == "f32" and opn == "add":
                            if rel_err(a, b) > FP32_TOL:
                                bad.append((n, m, kind, tname, opn, mode, mk is not None, rel_err(a, b)))
                        elif not np.array_equal(a, b):
                            bad.append((n, m, kind, tname, opn, mode, mk is not None))
    assert not bad, bad[:10]


def test_all_any_vs_reference_cuda(dr, R):
    for size in (1, 3, 4, 5, 1000, 1 << 20):
        for fill in (0, 1):
            for flip in (None, 0, size - 1, size // 2):
                m = np.full(size, fill, dtype=np.uint8)
                if flip is not None:
                    m[flip] ^= 1
                buf = torch.zeros(size + 64, dtype=torch.uint8, device="cuda")
                buf[:size].copy_(torch.from_numpy(m))
                ours = (dr.jit_all(CUDA, buf, size), dr.jit_any(CUDA, buf, size))
                buf[:size].copy_(torch.from_numpy(m))
                theirs = (ref_call(R, R.all, buf.data_ptr(), size),)
                buf[:size].copy_(torch.from_numpy(m))
                buf[size:].zero_()
                theirs += (ref_call(R, R.any, buf.data_ptr(), size),)
                assert ours == theirs, (size, fill, flip, ours, theirs)
