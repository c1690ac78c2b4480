"""GPU parity: compress and block_mkperm through the C-ABI vs the CPU oracle
(bit-exact: indices, counts, permutation, offsets records)."""
import ctypes

import numpy as np
import pytest

import oracle
from cases import cubic_sizes, fmix32, key_input, mask_input
from util import empty_dev, to_dev, to_host

pytestmark = pytest.mark.gpu
CUDA = 1


def run_compress(dr, m, offset=0):
    d_in = to_dev(m, offset)
    d_out = empty_dev(m.size, np.uint32)
    cnt = dr.jit_compress(CUDA, d_in, m.size, d_out)
    return to_host(d_out, np.uint32)[:cnt].copy(), cnt


def test_compress_grid(dr, O):
    # tests/reductions.cpp:269-313 (12_compress): sizes 23*i^3+1
    bad = []
    for size in cubic_sizes(30):
        for dens in (0.0, 0.01, 0.5, 0.99, 1.0):
            m = mask_input(size, dens)
            idx, cnt = run_compress(dr, m)
            ridx, rcnt = O.compress(m)
            if cnt != rcnt or not np.array_equal(idx, ridx):
                bad.append((size, dens, cnt, rcnt))
    assert not bad, bad


def test_compress_sparse_ones(dr, O):
    # the reference test places 23*j^3+1 ones at rand() positions
    rng = np.random.default_rng(0)
    bad = []
    for i in (1, 3, 8, 15, 29):
        size = 23 * i ** 3 + 1
        for j in (0, 1, i // 2, i):
            m = np.zeros(size, dtype=np.uint8)
            m[rng.integers(0, size, 23 * j ** 3 + 1)] = 1
            idx, cnt = run_compress(dr, m)
            ridx, rcnt = O.compress(m)
            if cnt != rcnt or not np.array_equal(idx, ridx):
                bad.append((size, j))
    assert not bad, bad


def test_compress_misaligned_and_edges(dr, O):
    bad = []
    for off in (1, 5, 15):
        for size in (1, 15, 16, 17, 4096, 16384, 16385, 100001):
            m = mask_input(size, 0.4, salt=off)
            idx, cnt = run_compress(dr, m, offset=off)
            ridx, rcnt = O.compress(m)
            if cnt != rcnt or not np.array_equal(idx, ridx):
                bad.append((off, size))
    # first / last element only
    for size in (1, 2, 16384, 16385, 1000000):
        for pos in (0, size - 1):
            m = np.zeros(size, dtype=np.uint8)
            m[pos] = 1
            idx, cnt = run_compress(dr, m)
            if cnt != 1 or idx[0] != pos:
                bad.append(("single", size, pos))
    assert dr.jit_compress(CUDA, 0, 0, 0) == 0  # size == 0 (src/cuda_ts.cpp:685-686)
    assert not bad, bad


def test_compress_misaligned_output_and_density_sweep(dr, O):
    # the dense rows of the stream kernel leave through 16-byte aligned bulk copies
    # with ragged ends: exercise every output alignment and densities on both sides
    # of the sparse / dense row threshold
    bad = []
    for dens in (0.05, 0.09, 0.12, 0.3, 0.7, 1.0):
        for off_in, off_out in ((0, 0), (0, 1), (3, 2), (7, 3)):
            size = 700001 + 4097 * off_out
            m = mask_input(size, dens, salt=off_out)
            d_in = to_dev(m, off_in)
            d_out = empty_dev(size + 8, np.uint32, off_out)
            d_out.fill_(-1)
            cnt = dr.jit_compress(CUDA, d_in, size, d_out)
            got = to_host(d_out, np.uint32)
            ridx, rcnt = O.compress(m)
            if cnt != rcnt or not np.array_equal(got[:cnt], ridx) or \
               not np.all(got[cnt:cnt + 8] == 0xffffffff):
                bad.append((dens, off_in, off_out, cnt, rcnt))
    assert not bad, bad


def test_compress_mixed_tile_kinds(dr, O):
    # tiles of 32768 entries: empty ones, sparse ones (<= 1024 set entries: expanded by the
    # pack kernel into the tile's own bit slot), tiles just above that threshold and dense
    # ones, in one mask; every input / output alignment; the last tile empty or sparse
    rng = np.random.default_rng(7)
    bad = []
    for ntiles, tail in ((37, 0), (64, 777), (130, 32767), (300, 5)):
        size = ntiles * 32768 + tail
        m = np.zeros(size, dtype=np.uint8)
        for t in range(ntiles + 1):
            lo, hi = t * 32768, min((t + 1) * 32768, size)
            if hi <= lo:
                continue
            kind = t % 6
            want = (0, 1, 1024, 1025, 3000, hi - lo)[kind]
            want = min(want, hi - lo)
            if t == ntiles:
                want = min(3, hi - lo) if ntiles % 2 else 0
            if want:
                m[lo + rng.choice(hi - lo, size=want, replace=False)] = 1
        ridx, rcnt = O.compress(m)
        for off_in, off_out in ((0, 0), (5, 1), (11, 3)):
            d_in = to_dev(m, off_in)
            d_out = empty_dev(size + 8, np.uint32, off_out)
            d_out.fill_(-1)
            cnt = dr.jit_compress(CUDA, d_in, size, d_out)
            got = to_host(d_out, np.uint32)
            if cnt != rcnt or not np.array_equal(got[:cnt], ridx) or \
               not np.all(got[cnt:cnt + 8] == 0xffffffff):
                bad.append((ntiles, tail, off_in, off_out, cnt, rcnt))
    assert not bad, bad


def test_compress_full_size(dr, O):
    # BASELINE.json configs[2]: 2^28-element mask at densities 0.01 / 0.5 / 0.99
    n = 1 << 28
    for dens in (0.01, 0.5, 0.99):
        m = mask_input(n, dens)
        idx, cnt = run_compress(dr, m)
        ridx, rcnt = O.compress(m)
        assert cnt == rcnt
        assert np.array_equal(idx, ridx)


def run_mkperm(dr, k, bs, buckets, want_offsets=True):
    import torch
    d_k = to_dev(k)
    d_perm = empty_dev(k.size, np.uint32)
    offsets = None
    if want_offsets:
        offsets = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
    uq = dr.jit_block_mkperm(CUDA, d_k, k.size, bs, buckets, d_perm, offsets)
    offs = offsets.numpy().view(np.uint32).copy() if want_offsets else None
    return to_host(d_perm, np.uint32), offs, uq


def test_mkperm_grid(dr, O):
    # tests/reductions.cpp:315-406 (13_mkperm) -- and stricter: the permutation is
    # stable, so it must equal the oracle's exactly, as must the offsets records
    # (ascending bucket id, like the reference's CPU path)
    bad = []
    for size in cubic_sizes(30)[::2]:
        for buckets in (1, 2, 16, 24, 300, 1024, 2048, 2049, 5000, 65536, 1000003):
            k = key_input(size, buckets)
            perm, offs, uq = run_mkperm(dr, k, size, buckets)
            rperm, roffs, ruq = O.block_mkperm(k, size, buckets)
            if uq != ruq or not np.array_equal(perm, rperm) or \
               not np.array_equal(offs[:4 * uq], roffs[:4 * ruq]) or offs[4 * buckets] != uq:
                bad.append((size, buckets, uq, ruq))
    assert not bad, bad


def test_mkperm_reference_test_semantics(dr):
    # exactly the checks of tests/reductions.cpp:357-400: buckets sorted by id,
    # per-bucket index sets equal to the sorted (key << 32 | index) list
    rng = np.random.default_rng(0)
    for size, buckets in ((24, 24), (14353, 185), (200000, 5000), (560948, 23 * 9 ** 3 + 1)):
        k = rng.integers(0, buckets, size).astype(np.uint32)
        perm, offs, uq = run_mkperm(dr, k, size, buckets)
        recs = offs[:4 * uq].reshape(-1, 4)
        recs = recs[np.argsort(recs[:, 0])]
        assert int(recs[:, 2].sum()) == size
        ref = np.sort((k.astype(np.uint64) << np.uint64(32)) | np.arange(size, dtype=np.uint64))
        pos = 0
        for bid, start, cnt, _ in recs:
            mine = np.sort(perm[start:start + cnt])
            chunk = ref[pos:pos + cnt]
            assert np.all((chunk >> np.uint64(32)) == bid)
            assert np.array_equal(mine, (chunk & np.uint64(0xFFFFFFFF)).astype(np.uint32))
            pos += cnt


def test_mkperm_blocked(dr, O):
    bad = []
    for size, bs, buckets in ((100000, 1000, 16), (100000, 12500, 300), (65536, 4096, 7),
                              (99999, 333, 40), (99999, 50000, 3000), (1000, 1, 5), (1000, 2, 5),
                              (250000, 33333, 100000),
                              # groups of at least half a tile: segmented ranked tiles
                              (100000, 4096, 16), (100000, 8192, 70), (100001, 8193, 5000), (300000, 4100, 3),
                              (1 << 20, 1 << 14, 1 << 16),
                              (1 << 20, 1 << 18, 16), (1500000, 400000, 1000), (3000001, 1 << 17, 70000)):
        k = key_input(size, buckets)
        perm, offs, uq = run_mkperm(dr, k, bs, buckets)
        rperm, _, ruq = O.block_mkperm(k, bs, buckets)
        if uq != ruq or not np.array_equal(perm, rperm):
            bad.append((size, bs, buckets))
        perm, _, uq = run_mkperm(dr, k, bs, buckets, want_offsets=False)
        if uq != 0 or not np.array_equal(perm, rperm):
            bad.append((size, bs, buckets, "no offsets"))
    assert not bad, bad


def test_mkperm_skewed_and_no_offsets(dr, O):
    for buckets in (16, 1024, 65536):
        k = key_input(1 << 20, buckets, skew=True)
        perm, offs, uq = run_mkperm(dr, k, k.size, buckets)
        rperm, roffs, ruq = O.block_mkperm(k, k.size, buckets)
        assert uq == ruq and np.array_equal(perm, rperm)
        assert np.array_equal(offs[:4 * uq], roffs[:4 * ruq])
        perm, _, uq0 = run_mkperm(dr, k, k.size, buckets, want_offsets=False)
        assert uq0 == 0 and np.array_equal(perm, rperm)


def test_mkperm_wide_keys_all_pass_forms(dr):
    # many digit passes: elements travel as (key, index) pairs first and as one
    # packed word once the remaining key bits fit beside the index; the stable
    # permutation is numpy's stable argsort.  Offsets are not requested (4 * B + 1
    # words would not fit for these bucket counts).
    rng = np.random.default_rng(7)
    for size, buckets in ((100000, 1 << 30), (100001, (1 << 24) + 1), (1 << 20, 1 << 18),
                          (8191, 0xffffffff), (300000, 70), (300000, 1 << 12), (300000, (1 << 13) + 5)):
        k = rng.integers(0, buckets, size, dtype=np.uint64).astype(np.uint32)
        k[:7] = buckets - 1
        for off in (0, 1):
            d_k = to_dev(k, off)
            d_perm = empty_dev(size, np.uint32, off)
            uq = dr.jit_block_mkperm(CUDA, d_k, size, size, buckets, d_perm, None)
            assert uq == 0
            assert np.array_equal(to_host(d_perm, np.uint32),
                                  np.argsort(k, kind="stable").astype(np.uint32)), (size, buckets, off)


def test_mkperm_histogram(dr):
    for size, buckets in ((1000, 7), (1 << 20, 1024), (1 << 20, 65536), (3000001, 100000), (1 << 20, 300000)):
        k = key_input(size, buckets)
        h = empty_dev(buckets, np.uint32)
        dr.mkperm_histogram(to_dev(k), size, buckets, h)
        assert np.array_equal(to_host(h, np.uint32), np.bincount(k, minlength=buckets).astype(np.uint32))


def test_mkperm_full_size(dr, O):
    # BASELINE.json configs[3]: 2^26 callee ids into 16 / 1024 / 65536 buckets
    n = 1 << 26
    for buckets in (16, 1024, 65536):
        k = key_input(n, buckets)
        perm, offs, uq = run_mkperm(dr, k, n, buckets)
        rperm, roffs, ruq = O.block_mkperm(k, n, buckets)
        assert uq == ruq
        assert np.array_equal(offs[:4 * uq], roffs[:4 * ruq])
        assert np.array_equal(perm, rperm)


def test_compress_nonbinary_mask_bytes(dr):
    # jit.h:2377-2379 defines only 0 / 1 mask bytes; every non-zero byte selects the
    # entry on BOTH code paths of jit_compress (single pass <= 32768 < bit-packed tiles)
    from cases import fmix32
    bad = []
    for size in (1000, 32768, 32769, 100003, (1 << 20) + 5):
        h = fmix32(np.arange(size, dtype=np.uint32))
        m = np.where((h & np.uint32(3)) == 0, np.uint8(0), (h >> np.uint32(8)).astype(np.uint8))
        # bytes like 0x02, 0x80, 0xfe appear; a few become 0 by chance -- that is the definition
        expect = np.flatnonzero(m).astype(np.uint32)
        idx, cnt = run_compress(dr, m)
        if cnt != expect.size or not np.array_equal(idx, expect):
            bad.append((size, cnt, expect.size))
    assert not bad, bad


def test_call_reduce_records_by_size(dr, O):
    # the mkperm step of jitc_var_call_reduce (src/call.cpp:1268-1389): ids in [0, id_bound]
    # (0 = null callable), bucket_count = id_bound + 1 (call.cpp:1292), records sorted by
    # size, largest first (call.cpp:1346-1352) -- ties by ascending id here.  The
    # permutation equals jit_block_mkperm's; the records are the oracle's, re-sorted.
    import torch
    bad = []
    for size in (1, 31, 1000, 40000, 300007):
        for id_bound in (1, 3, 15, 100, 1023, 5000, 70000):
            ids = key_input(size, id_bound + 1)
            if size > 100:  # skew: a dominant callable and a rare null bucket
                ids = np.where(fmix32(np.arange(size, dtype=np.uint32) ^ np.uint32(77)) % np.uint32(10) < 6,
                               np.uint32(min(2, id_bound)), ids).astype(np.uint32)
            buckets = id_bound + 1
            d_perm = empty_dev(size, np.uint32)
            offsets = torch.zeros(4 * buckets + 1, dtype=torch.int32).pin_memory()
            uq = dr.call_reduce(to_dev(ids), size, id_bound, d_perm, offsets)
            rperm, roffs, ruq = O.block_mkperm(ids, size, buckets)
            rec = roffs[:4 * ruq].reshape(-1, 4)
            order = np.lexsort((rec[:, 0], -rec[:, 2].astype(np.int64)))
            got = offsets.numpy().view(np.uint32)
            if uq != ruq or got[4 * buckets] != uq or not np.array_equal(to_host(d_perm, np.uint32), rperm) or \
               not np.array_equal(got[:4 * uq].reshape(-1, 4), rec[order]):
                bad.append((size, id_bound, uq, ruq))
    assert not bad, bad
