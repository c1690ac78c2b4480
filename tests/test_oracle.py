"""CPU-only: pin the oracle (oracle/oracle.c) against
  (a) the committed fixtures generated from the reference's own CPU
      implementation (tests/golden/reference_cpu.json, make_golden.py),
  (b) the closed forms the reference's tests use (tests/reductions.cpp:15-70),
  (c) the reference CPU implementation live, when oracle/_ref has been built.
"""
import numpy as np
import pytest

import oracle
from cases import (RED_SIZES, cubic_sizes, fmix32, key_input, mask_input, red_pairs,
                   u32_input, u64_input)
from util import sha1

VT, OP = oracle.VT, oracle.OP


def test_fmix32_matches_c(O):
    idx = np.array([0, 1, 2, 1000, 2 ** 31, 2 ** 32 - 1], dtype=np.uint32)
    assert [int(v) for v in fmix32(idx)] == [O.lib.oracle_fmix32(int(i)) for i in idx]


def test_identities_golden(O, golden):
    for vt in (4, 7, 8, 9, 10, 13, 14, 15):
        for op in range(1, 7):
            assert O.reduce_identity(vt, op) == golden[f"identity/{vt}/{op}"]


@pytest.mark.parametrize("tname", ["u32", "u64"])
def test_block_reduce_grid_golden(O, golden, tname):
    # tests/reductions.cpp:109-151 (02_block_reduce_u32, 03_block_reduce_u64)
    for size, bs in red_pairs():
        x = u32_input(size) if tname == "u32" else u64_input(size)
        r = O.block_reduce(VT[tname], OP["add"], x, bs)
        assert sha1(r) == golden[f"block_reduce/{tname}/add/{size}/{bs}"]["sha1"], (size, bs)


@pytest.mark.parametrize("tname", ["u32", "u64"])
def test_prefix_grid_golden(O, golden, tname):
    # tests/reductions.cpp:153-267 (04..11_block_prefix_reduce_*)
    for size, bs in red_pairs(max_size=170000):
        x = u32_input(size) if tname == "u32" else u64_input(size)
        for excl in (0, 1):
            for rev in (0, 1):
                r = O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev)
                key = f"prefix/{tname}/add/{size}/{bs}/{excl}/{rev}"
                assert sha1(r) == golden[key]["sha1"], key


def test_prefix_largest_size_golden(O, golden):
    size = RED_SIZES[-1]
    x = u32_input(size)
    for bs in (1, 7, 1024, 169541, size):
        for excl, rev in ((0, 0), (1, 1)):
            r = O.block_prefix_reduce(VT["u32"], OP["add"], x, bs, excl, rev)
            assert sha1(r) == golden[f"prefix/u32/add/{size}/{bs}/{excl}/{rev}"]["sha1"]


def test_other_ops_golden(O, golden):
    for size, bs in ((1000, 7), (1000, 1000), (163880, 333), (163880, 163880), (4097, 64)):
        x32 = u32_input(size)
        arrays = {
            "u32": x32, "i32": x32.view(np.int32),
            "u64": (x32.astype(np.uint64) << np.uint64(17)) ^ x32.astype(np.uint64),
            "i64": ((x32.astype(np.uint64) << np.uint64(33)) ^ x32.astype(np.uint64)).view(np.int64),
        }
        for tname, x in arrays.items():
            for opn in ("add", "mul", "min", "max", "and_", "or_"):
                r = O.block_reduce(VT[tname], OP[opn], x, bs)
                assert sha1(r) == golden[f"ops_reduce/{tname}/{opn}/{size}/{bs}"]["sha1"]
                r = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, 1, 0)
                assert sha1(r) == golden[f"ops_prefix/{tname}/{opn}/{size}/{bs}/1/0"]["sha1"]
                r = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, 0, 1)
                assert sha1(r) == golden[f"ops_prefix/{tname}/{opn}/{size}/{bs}/0/1"]["sha1"]


def test_block_reduce_const_closed_form(O):
    # tests/reductions.cpp:38-46 (block_sum_ref_const): summing ones
    for size, bs in red_pairs(max_size=170000):
        r = O.block_reduce(VT["u32"], OP["add"], np.ones(size, dtype=np.uint32), bs)
        blocks = (size + bs - 1) // bs
        expect = np.minimum(size - np.arange(blocks, dtype=np.int64) * bs, bs).astype(np.uint32)
        assert np.array_equal(r, expect)


def test_compress_golden(O, golden):
    # tests/reductions.cpp:269-313
    for size in cubic_sizes(30):
        for dens in (0.0, 0.01, 0.5, 0.99, 1.0):
            m = mask_input(size, dens)
            idx, cnt = O.compress(m)
            g = golden[f"compress/{size}/{dens}"]
            assert cnt == g["count"] and sha1(idx) == g["sha1"]
            assert np.array_equal(idx, np.nonzero(m)[0].astype(np.uint32))


def test_mkperm_golden(O, golden):
    # tests/reductions.cpp:315-406
    for size in cubic_sizes(30)[::3]:
        for buckets in (1, 2, 16, 24, 1024, 5000, 65536):
            k = key_input(size, buckets)
            perm, offs, uq = O.block_mkperm(k, size, buckets)
            g = golden[f"mkperm/{size}/{buckets}"]
            assert uq == g["unique"] and sha1(perm) == g["sha1"]
            assert sha1(offs[:4 * uq]) == g["offsets_sha1"]
            assert np.array_equal(perm, np.argsort(k, kind="stable").astype(np.uint32))


def test_mkperm_blocked_golden(O, golden):
    for size, bs, buckets in ((100000, 1000, 16), (100000, 12500, 300), (65536, 4096, 7),
                              (99999, 333, 40)):
        k = key_input(size, buckets)
        perm, _, uq = O.block_mkperm(k, bs, buckets)
        g = golden[f"mkperm_blocked/{size}/{bs}/{buckets}"]
        assert uq == g["unique"] == 0 and sha1(perm) == g["sha1"]


def test_allany_golden(O, golden):
    for size in (1, 3, 4, 5, 1000, 4099):
        f = np.zeros(size, dtype=np.uint8)
        t = np.ones(size, dtype=np.uint8)
        g = golden[f"allany/{size}"]
        assert (O.all(t), O.any(t), O.all(f), O.any(f)) == (g["all_t"], g["any_t"], g["all_f"], g["any_f"])
        f[size // 2] = 1
        t[size // 2] = 0
        g = golden[f"allany_flip/{size}"]
        assert (O.all(t), O.any(t), O.all(f), O.any(f)) == (g["all_t"], g["any_t"], g["all_f"], g["any_f"])


def test_errors(O):
    x = np.ones(10, dtype=np.uint32)
    with pytest.raises(ValueError):
        O.block_reduce(VT["u32"], OP["add"], x, 0)
    with pytest.raises(ValueError):
        O.block_reduce(VT["u32"], OP["add"], x, 11)
    with pytest.raises(ValueError):
        O.block_prefix_reduce(VT["f32"], OP["and_"], x.astype(np.float32), 2, 0, 0)
    with pytest.raises(ValueError):
        O.scatter_reduce(VT["u32"], OP["mul"], x, x, x)


def test_float_paths_against_fp64(O):
    rng = np.random.default_rng(1)
    x = rng.random(100003).astype(np.float32)
    for bs in (3, 1000, 100003):
        wide = O.block_reduce(VT["f32"], OP["add"], x, bs, wide=True)
        blocks = (x.size + bs - 1) // bs
        expect = np.add.reduceat(x.astype(np.float64), np.arange(blocks) * bs)
        assert np.allclose(wide, expect, rtol=1e-6)
        pw = O.block_prefix_reduce(VT["f32"], OP["add"], x, bs, 0, 0, wide=True)
        assert np.isclose(pw[bs - 1], expect[0], rtol=1e-6)
    h = rng.standard_normal(999).astype(np.float16)
    r = O.block_reduce(VT["f16"], OP["add"], h, 999)
    assert np.isclose(float(r[0]), h.astype(np.float64).sum(), atol=0.05)
    assert O.block_reduce(VT["f16"], OP["max"], h, 999)[0] == h.max()
    d = O.reduce_dot(VT["f32"], x, x, wide=True)
    assert np.isclose(float(d[0]), np.dot(x.astype(np.float64), x.astype(np.float64)), rtol=1e-6)


def test_scatter_oracle_semantics(O):
    rng = np.random.default_rng(2)
    n, m = 5000, 97
    idx = rng.integers(0, m, n).astype(np.uint32)
    val = rng.integers(0, 1000, n).astype(np.uint32)
    mask = (rng.random(n) < 0.7).astype(np.uint8)
    tgt = np.zeros(m, dtype=np.uint32)
    r = O.scatter_reduce(VT["u32"], OP["add"], tgt, val, idx, mask)
    expect = np.bincount(idx[mask != 0], weights=val[mask != 0], minlength=m).astype(np.uint32)
    assert np.array_equal(r, expect)
    r = O.scatter_reduce(VT["u32"], OP["max"], tgt, val, idx)
    e2 = np.zeros(m, dtype=np.uint32)
    np.maximum.at(e2, idx, val)
    assert np.array_equal(r, e2)
    # tests/mem.cpp:137-174 (10_scatter_atomic_rmw): 16 adds of 1.0 with duplicates
    target = np.zeros(5, dtype=np.float32)
    index = np.array([0, 0, 1, 2, 2, 2, 3, 4, 4, 4, 4, 0, 1, 2, 3, 4], dtype=np.uint32)
    r = O.scatter_reduce(VT["f32"], OP["add"], target, np.ones(16, dtype=np.float32), index)
    assert np.array_equal(r, np.bincount(index, minlength=5).astype(np.float32))


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_against_reference_cpu_live(O):
    R = oracle.Reference()
    for size, bs in ((333, 7), (16384, 1024), (169541, 333), (169541, 169541)):
        x = u32_input(size)
        assert np.array_equal(O.block_reduce(VT["u32"], OP["add"], x, bs),
                              R.block_reduce(VT["u32"], OP["add"], x, bs))
        for excl in (0, 1):
            for rev in (0, 1):
                assert np.array_equal(
                    O.block_prefix_reduce(VT["u32"], OP["max"], x, bs, excl, rev),
                    R.block_prefix_reduce(VT["u32"], OP["max"], x, bs, excl, rev))
    m = mask_input(100001, 0.3)
    a, ca = O.compress(m)
    b, cb = R.compress(m)
    assert ca == cb and np.array_equal(a, b)
    k = key_input(100001, 777)
    pa, oa, ua = O.block_mkperm(k, 100001, 777)
    pb, ob, ub = R.block_mkperm(k, 100001, 777)
    assert ua == ub and np.array_equal(pa, pb) and np.array_equal(oa[:4 * ua], ob[:4 * ub])
