"""GPU parity: blocked prefix reductions through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

import oracle
from cases import RED_SIZES, f32_input, fmix32, red_pairs, u32_input, u64_input
from util import empty_dev, rel_err, to_dev, to_host

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
CUDA = 1


def run_scan(dr, vt, op, x, bs, excl, rev, offset=0, inplace=False):
    d_in = to_dev(x, offset)
    d_out = d_in if inplace else empty_dev(x.size, x.dtype, offset)
    dr.jit_block_prefix_reduce(CUDA, vt, op, x.size, bs, excl, rev, d_in, d_out)
    return to_host(d_out, x.dtype)


@pytest.mark.parametrize("tname", ["u32", "u64"])
@pytest.mark.parametrize("excl", [0, 1])
@pytest.mark.parametrize("rev", [0, 1])
def test_prefix_grid(dr, O, tname, excl, rev):
    # tests/reductions.cpp:153-267 (04..11) on the reference's size grid
    bad = []
    for size, bs in red_pairs():
        if size == RED_SIZES[-1] and bs not in (1, 2, 7, 333, 1024, 16384, 169541, size):
            continue
        x = u32_input(size) if tname == "u32" else u64_input(size)
        got = run_scan(dr, VT[tname], OP["add"], x, bs, excl, rev)
        if not np.array_equal(got, O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev)):
            bad.append((size, bs))
    assert not bad, bad


def int_input(tname, size):
    h = u32_input(size)
    if tname == "u32":
        return h
    if tname == "i32":
        return h.view(np.int32)
    w = (h.astype(np.uint64) << np.uint64(29)) ^ fmix32(h).astype(np.uint64)
    return w if tname == "u64" else w.view(np.int64)


@pytest.mark.parametrize("tname", ["u32", "i32", "u64", "i64"])
def test_prefix_all_int_ops(dr, O, tname):
    bad = []
    for size, bs in ((1, 1), (9, 4), (1000, 7), (4100, 64), (70001, 333), (70001, 4096),
                     (70001, 5000), (70001, 70001), (300000, 100000)):
        x = int_input(tname, size)
        for opn in ("add", "mul", "min", "max", "and_", "or_"):
            for excl, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
                got = run_scan(dr, VT[tname], OP[opn], x, bs, excl, rev)
                ref = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, excl, rev)
                if not np.array_equal(got, ref):
                    bad.append((size, bs, opn, excl, rev))
    assert not bad, bad


def test_prefix_pow2_blocks(dr, O):
    bad = []
    for size in (1 << 16, (1 << 16) + 37, 1000001):
        x = u32_input(size)
        for lg in range(0, 18):
            bs = 1 << lg
            if bs > size:
                continue
            for excl, rev in ((0, 0), (1, 0), (1, 1)):
                got = run_scan(dr, VT["u32"], OP["add"], x, bs, excl, rev)
                if not np.array_equal(got, O.block_prefix_reduce(VT["u32"], OP["add"], x, bs, excl, rev)):
                    bad.append((size, bs, excl, rev))
    assert not bad, bad


def test_prefix_inplace_and_misaligned(dr, O):
    bad = []
    for size, bs in ((1000, 7), (100003, 4096), (100003, 100003), (100003, 16)):
        x = u32_input(size)
        for excl, rev in ((0, 0), (1, 1)):
            ref = O.block_prefix_reduce(VT["u32"], OP["add"], x, bs, excl, rev)
            if not np.array_equal(run_scan(dr, VT["u32"], OP["add"], x, bs, excl, rev, inplace=True), ref):
                bad.append(("inplace", size, bs, excl, rev))
            for off in (1, 3):
                if not np.array_equal(run_scan(dr, VT["u32"], OP["add"], x, bs, excl, rev, offset=off), ref):
                    bad.append(("offset", off, size, bs, excl, rev))
    assert not bad, bad


# fp32 Add: every prefix within 1e-5 relative of the fp64-accumulated oracle
# (all-positive inputs).  f64: 1e-12.  f16: 1 ulp.
def test_prefix_float(dr, O):
    bad = []
    for tname, tol in (("f32", 1e-5), ("f64", 1e-12)):
        dt = oracle.NP_OF_VT[VT[tname]]
        for size, bs in ((1000, 7), (65536, 256), (100003, 4096), (100003, 100003), (1 << 21, 1 << 21)):
            x = f32_input(size).astype(dt)
            for excl, rev in ((0, 0), (1, 1)):
                got = run_scan(dr, VT[tname], OP["add"], x, bs, excl, rev)
                ref = O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev, wide=True)
                sel = ref != 0
                err = rel_err(got[sel], ref[sel])
                if err > tol or not np.all(got[~sel] == 0):
                    bad.append((tname, size, bs, excl, rev, err))
            got = run_scan(dr, VT[tname], OP["max"], x, bs, 0, 0)
            if not np.array_equal(got, O.block_prefix_reduce(VT[tname], OP["max"], x, bs, 0, 0)):
                bad.append((tname, size, bs, "max"))
    x = (f32_input(20000) * 0.01).astype(np.float16)
    got = run_scan(dr, VT["f16"], OP["add"], x, 5000, 0, 0).astype(np.float64)
    ref = O.block_prefix_reduce(VT["f16"], OP["add"], x, 5000, 0, 0, wide=True).astype(np.float64)
    if rel_err(got[ref != 0], ref[ref != 0]) > 2.0 ** -10:
        bad.append(("f16", rel_err(got[ref != 0], ref[ref != 0])))
    assert not bad, bad


def test_prefix_block_size_one_and_errors(dr, O):
    x = u32_input(1000)
    assert np.array_equal(run_scan(dr, VT["u32"], OP["add"], x, 1, 0, 0), x)
    assert np.array_equal(run_scan(dr, VT["u32"], OP["add"], x, 1, 1, 0), np.zeros_like(x))
    got = run_scan(dr, VT["u32"], OP["and_"], x, 1, 1, 0)
    assert np.all(got == 0xFFFFFFFF)
    x64 = x.astype(np.int64).view(np.int64)
    got = run_scan(dr, VT["i64"], OP["min"], x64, 1, 1, 0)  # 8-byte fill pattern
    assert np.all(got == np.iinfo(np.int64).max)
    f = f32_input(100)
    got = run_scan(dr, VT["f32"], OP["mul"], f, 1, 1, 0)
    assert np.all(got == 1.0)
    d = to_dev(x)
    with pytest.raises(RuntimeError, match="invalid block size"):
        dr.jit_block_prefix_reduce(CUDA, VT["u32"], OP["add"], 1000, 0, 0, 0, d, d)
    with pytest.raises(RuntimeError, match="invalid block size"):
        dr.jit_block_prefix_reduce(CUDA, VT["u32"], OP["add"], 1000, 1001, 0, 0, d, d)
    with pytest.raises(RuntimeError, match="no existing kernel"):
        dr.jit_block_prefix_reduce(CUDA, VT["f32"], OP["or_"], 1000, 10, 0, 0, d, d)


def test_prefix_carry_api(dr, O):
    import torch
    bad = []
    for size in (1, 100, 4096, 100003):
        x = u32_input(size)
        for excl, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
            carry = to_dev(np.array([12345], dtype=np.uint32))
            cout = empty_dev(1, np.uint32)
            d_out = empty_dev(size, np.uint32)
            dr.prefix_reduce_carry(VT["u32"], OP["add"], size, excl, rev, to_dev(x), d_out, carry, cout)
            ref = (O.block_prefix_reduce(VT["u32"], OP["add"], x, size, excl, rev) + np.uint32(12345)).astype(np.uint32)
            total = np.uint32(int(x.sum(dtype=np.uint64) + 12345) & 0xFFFFFFFF)
            if not np.array_equal(to_host(d_out, np.uint32), ref) or to_host(cout, np.uint32)[0] != total:
                bad.append((size, excl, rev))
            # without carry-in
            dr.prefix_reduce_carry(VT["u32"], OP["add"], size, excl, rev, to_dev(x), d_out, None, cout)
            ref = O.block_prefix_reduce(VT["u32"], OP["add"], x, size, excl, rev)
            if not np.array_equal(to_host(d_out, np.uint32), ref):
                bad.append((size, excl, rev, "nocarry"))
    assert not bad, bad


def test_prefix_seeded_api(dr, O):
    # b200_prefix_reduce_seeded: the caller supplies the exclusive prefix of every
    # 32 KiB tile (what the sharded front end derives from its reduce pass); the
    # result must equal the whole-array scan with a carry -- bit-exact for integers
    bad = []
    for tname, dt in (("u32", np.uint32), ("u64", np.uint64), ("f32", np.float32)):
        tile = dr.scan_tile_elems(VT[tname])
        assert tile * np.dtype(dt).itemsize == 32768
        for size in (1, tile - 1, tile, 5 * tile + 17, (1 << 22) + 12345):
            x = {"u32": u32_input, "u64": u64_input, "f32": f32_input}[tname](size)
            ntiles = -(-size // tile)
            pad = np.zeros(ntiles * tile, dtype=dt)
            pad[:size] = x
            tsums = pad.reshape(ntiles, tile).sum(axis=1, dtype=np.float64 if tname == "f32" else dt)
            carry = dt(12345)
            for excl, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
                if rev:
                    seeds = (np.cumsum(tsums[::-1])[::-1] - tsums + carry).astype(dt)
                else:
                    seeds = (np.cumsum(tsums) - tsums + carry).astype(dt)
                d_out = empty_dev(size, dt)
                dr.prefix_reduce_seeded(VT[tname], OP["add"], size, excl, rev, to_dev(x), d_out, to_dev(seeds))
                got = to_host(d_out, dt)
                if tname == "f32":
                    x64 = x.astype(np.float64)
                    inc = np.cumsum(x64[::-1])[::-1] if rev else np.cumsum(x64)
                    ref = inc - (x64 if excl else 0) + float(carry)
                    if np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) > 1e-5:  # fp32 Add: 1e-5
                        bad.append((tname, size, excl, rev))
                else:
                    ref = (O.block_prefix_reduce(VT[tname], OP["add"], x, size, excl, rev) + carry).astype(dt)
                    if not np.array_equal(got, ref):
                        bad.append((tname, size, excl, rev))
    assert not bad, bad
    with pytest.raises(RuntimeError):  # misaligned arrays are not served by the seeded kernel
        dr.prefix_reduce_seeded(VT["u32"], OP["add"], 100000, 0, 0, to_dev(u32_input(100000), 1),
                                empty_dev(100000, np.uint32, 1), to_dev(np.zeros(16, dtype=np.uint32)))


def test_full_size_scan(dr, O):
    # 2^28 u32 exclusive prefix sum (BASELINE.json configs[0]), bit-exact
    n = 1 << 28
    x = u32_input(n)
    d = to_dev(x)
    dr.jit_block_prefix_reduce(CUDA, VT["u32"], OP["add"], n, n, 1, 0, d, d)  # in place
    got = to_host(d, np.uint32)
    ref = O.block_prefix_reduce(VT["u32"], OP["add"], x, n, 1, 0)
    assert np.array_equal(got, ref)
    del ref
    # blocked, fp32: last element of every inclusive block prefix == block sum
    xf = f32_input(n)
    df = to_dev(xf)
    out = empty_dev(n, np.float32)
    for bs in (2, 64, 4096):
        dr.jit_block_prefix_reduce(CUDA, VT["f32"], OP["add"], n, bs, 0, 0, df, out)
        got = to_host(out, np.float32)
        sums = xf.reshape(-1, bs).astype(np.float64).sum(axis=1)
        assert rel_err(got.reshape(-1, bs)[:, -1], sums) <= 1e-5


@pytest.mark.parametrize("tname", ["u32", "i32", "u64", "i64"])
def test_prefix_streaming_paths(dr, O, tname):
    """The persistent bulk-copy kernels of scan_fast.cu: sizes beyond one wave of
    CTAs (several tiles per CTA), power-of-two blocks inside a tile (POW2), whole
    arrays and tile-multiple power-of-two blocks (CHAIN), ragged tails, reverse,
    in place.  Integer results are bit-exact."""
    bad = []
    sizes = [(1 << 22) + 12345, 3 * (1 << 20), 2500001]
    for size in sizes:
        x = int_input(tname, size)
        for bs in (2, 4, 8, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18,
                   1 << 20, size):
            for opn, excl, rev in (("add", 0, 0), ("add", 1, 1), ("max", 1, 0), ("min", 0, 1)):
                got = run_scan(dr, VT[tname], OP[opn], x, bs, excl, rev, inplace=(bs == 64))
                ref = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, excl, rev)
                if not np.array_equal(got, ref):
                    bad.append((size, bs, opn, excl, rev))
    assert not bad, bad


@pytest.mark.parametrize("tname", ["u32", "u64"])
def test_prefix_streaming_block_groups(dr, O, tname):
    """Blocks of 2 .. 16 tiles are scanned by one CTA each, tile after tile, without a global look-back
    (scan_fast.cu, `log2_group`); larger blocks and ragged reverse scans keep the per-tile chain.  Sizes of
    many groups per CTA, whole and ragged; bit-exact."""
    bad = []
    for size in (1 << 25, (1 << 25) + (1 << 16) + 77, 40000003):
        x = int_input(tname, size)
        for bs in (1 << 13, 1 << 14, 1 << 16, 1 << 17, 1 << 18):
            for excl, rev in ((1, 0), (0, 1)):
                got = run_scan(dr, VT[tname], OP["add"], x, bs, excl, rev)
                ref = O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev)
                if not np.array_equal(got, ref):
                    bad.append((size, bs, excl, rev))
    assert not bad, bad


@pytest.mark.parametrize("tname", ["u32", "i32", "u64", "i64"])
def test_prefix_two_stream_chain(dr, O, tname):
    """scan_ahead_kernel (scan_fast.cu): whole arrays and blocks of more than 16 tiles, from 1024 tiles on --
    a reduce stream ahead of the scan stream, one prefix stream over the tile aggregates.  Ragged ends (reverse:
    the block grid is out of phase with the tiles, tile 0 is not a block start), in place, every operation,
    blocks of few tiles with a ragged reverse end (no groups).  Bit-exact."""
    bad = []
    tile = 32768 // np.dtype(int_input(tname, 1).dtype).itemsize
    for size in (1024 * tile, 1500 * tile + 12345, 3000 * tile + 1):
        x = int_input(tname, size)
        for bs in (size, 32 * tile, 64 * tile, 2 * tile):
            for opn, excl, rev in (("add", 1, 0), ("add", 0, 1), ("max", 1, 1), ("min", 0, 0), ("or_", 1, 1)):
                if opn == "or_" and tname in ("i32", "i64"):
                    continue
                got = run_scan(dr, VT[tname], OP[opn], x, bs, excl, rev, inplace=(opn == "max"))
                ref = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, excl, rev)
                if not np.array_equal(got, ref):
                    bad.append((size, bs, opn, excl, rev))
    assert not bad, bad


def test_prefix_two_stream_chain_carry_and_float(dr, O):
    bad = []
    size = 1100 * 8192 + 77
    x = u32_input(size)
    for excl, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
        carry = to_dev(np.array([12345], dtype=np.uint32))
        cout = empty_dev(1, np.uint32)
        d_out = empty_dev(size, np.uint32)
        dr.prefix_reduce_carry(VT["u32"], OP["add"], size, excl, rev, to_dev(x), d_out, carry, cout)
        ref = (O.block_prefix_reduce(VT["u32"], OP["add"], x, size, excl, rev) + np.uint32(12345)).astype(np.uint32)
        total = np.uint32(int(x.sum(dtype=np.uint64) + 12345) & 0xFFFFFFFF)
        if not np.array_equal(to_host(d_out, np.uint32), ref) or to_host(cout, np.uint32)[0] != total:
            bad.append((excl, rev))
    assert not bad, bad
    # fp32 / fp64 / fp16: every prefix against an fp64 accumulation
    xf = f32_input(size)
    for dt, vt, tol in ((np.float32, "f32", 1e-5), (np.float64, "f64", 1e-12)):
        xs = xf.astype(dt)
        for excl, rev in ((1, 0), (0, 1)):
            got = run_scan(dr, VT[vt], OP["add"], xs, size, excl, rev).astype(np.float64)
            v = xs.astype(np.float64)[::-1] if rev else xs.astype(np.float64)
            ref = np.cumsum(v)
            if excl:
                ref = np.concatenate(([0.0], ref[:-1]))
            if rev:
                ref = ref[::-1]
            assert rel_err(got, ref) <= tol, (vt, excl, rev, rel_err(got, ref))
        # the two-stream scan sums in a fixed order: run-to-run identical
        a = run_scan(dr, VT[vt], OP["add"], xs, size, 1, 0)
        b = run_scan(dr, VT[vt], OP["add"], xs, size, 1, 0)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("tname", ["u32", "i64"])
def test_prefix_streaming_odd_blocks(dr, O, tname):
    """SEG mode of scan_fast.cu: block sizes that are not powers of two -- smaller than a
    vector, around a tile (8192 / 4096 elements), multiples of a tile, larger than many
    tiles -- at sizes of several tiles per CTA, ragged tails, every direction, in place."""
    bad = []
    for size in ((1 << 22) + 12345, 2500001):
        x = int_input(tname, size)
        for bs in (3, 5, 6, 100, 1000, 4095, 4097, 8191, 8193, 3 * 8192, 100000, 1234567, size - 1):
            for opn, excl, rev in (("add", 0, 0), ("add", 1, 0), ("add", 1, 1), ("max", 0, 1)):
                got = run_scan(dr, VT[tname], OP[opn], x, bs, excl, rev, inplace=(bs == 100))
                ref = O.block_prefix_reduce(VT[tname], OP[opn], x, bs, excl, rev)
                if not np.array_equal(got, ref):
                    bad.append((size, bs, opn, excl, rev, int(np.flatnonzero(got != ref)[0])))
    assert not bad, bad


def test_prefix_streaming_odd_blocks_float(dr, O):
    bad = []
    size = (1 << 21) + 77
    for tname, tol in (("f32", 1e-5), ("f64", 1e-12)):
        dt = oracle.NP_OF_VT[VT[tname]]
        x = f32_input(size).astype(dt)
        for bs in (3, 100, 1000, 8193, 100000):
            for excl, rev in ((0, 0), (1, 0), (1, 1)):
                got = run_scan(dr, VT[tname], OP["add"], x, bs, excl, rev)
                ref = O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev, wide=True)
                sel = ref != 0
                err = rel_err(got[sel], ref[sel])
                if err > tol or not np.all(got[~sel] == 0):
                    bad.append((tname, bs, excl, rev, err))
            got = run_scan(dr, VT[tname], OP["min"], x, bs, 0, 1)
            if not np.array_equal(got, O.block_prefix_reduce(VT[tname], OP["min"], x, bs, 0, 1)):
                bad.append((tname, bs, "min"))
    x = (f32_input(1 << 18) * 0.01).astype(np.float16)
    for bs in (3, 1000, 20000):
        got = run_scan(dr, VT["f16"], OP["add"], x, bs, 0, 0).astype(np.float64)
        ref = O.block_prefix_reduce(VT["f16"], OP["add"], x, bs, 0, 0, wide=True).astype(np.float64)
        sel = ref != 0
        if rel_err(got[sel], ref[sel]) > 2.0 ** -10:
            bad.append(("f16", bs, rel_err(got[sel], ref[sel])))
        got = run_scan(dr, VT["f16"], OP["max"], x, bs, 1, 1)
        if not np.array_equal(got, O.block_prefix_reduce(VT["f16"], OP["max"], x, bs, 1, 1)):
            bad.append(("f16", bs, "max"))
    assert not bad, bad


def test_prefix_streaming_float_types(dr, O):
    # f32 / f64 / f16 through the streaming kernels; tolerances as in test_prefix_float
    bad = []
    size = (1 << 21) + 77
    for tname, tol in (("f32", 1e-5), ("f64", 1e-12)):
        dt = oracle.NP_OF_VT[VT[tname]]
        x = f32_input(size).astype(dt)
        for bs in (2, 16, 128, 1024, 4096, 1 << 15, size):
            for excl, rev in ((0, 0), (1, 0), (1, 1)):
                got = run_scan(dr, VT[tname], OP["add"], x, bs, excl, rev)
                ref = O.block_prefix_reduce(VT[tname], OP["add"], x, bs, excl, rev, wide=True)
                sel = ref != 0
                err = rel_err(got[sel], ref[sel])
                if err > tol or not np.all(got[~sel] == 0):
                    bad.append((tname, bs, excl, rev, err))
            got = run_scan(dr, VT[tname], OP["min"], x, bs, 0, 1)
            if not np.array_equal(got, O.block_prefix_reduce(VT[tname], OP["min"], x, bs, 0, 1)):
                bad.append((tname, bs, "min"))
    x = (f32_input(1 << 18) * 0.01).astype(np.float16)
    for bs in (8, 256, 8192):
        got = run_scan(dr, VT["f16"], OP["add"], x, bs, 0, 0).astype(np.float64)
        ref = O.block_prefix_reduce(VT["f16"], OP["add"], x, bs, 0, 0, wide=True).astype(np.float64)
        sel = ref != 0
        if rel_err(got[sel], ref[sel]) > 2.0 ** -10:
            bad.append(("f16", bs, rel_err(got[sel], ref[sel])))
        got = run_scan(dr, VT["f16"], OP["max"], x, bs, 0, 0)
        if not np.array_equal(got, O.block_prefix_reduce(VT["f16"], OP["max"], x, bs, 0, 0)):
            bad.append(("f16", bs, "max"))
    assert not bad, bad
