"""Deterministic inputs shared by the golden-vector generator and the tests.

Inputs are counter based (fmix32 of the index, tests/reductions.cpp:5-13 of the
reference) so that every consumer regenerates identical data without an RNG.
"""
import numpy as np


def fmix32(i):
    h = (np.asarray(i, dtype=np.uint32) + np.uint32(1)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85ebca6b)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xc2b2ae35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


# tests/reductions.cpp:73-76 (used for both size and block_size)
RED_SIZES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 32, 60, 128, 250, 333, 1024, 16384,
             16388 * 10, 9973 * 17, 98973 * 17 * 3]


def red_pairs(max_size=None):
    for size in RED_SIZES:
        if max_size and size > max_size:
            continue
        for bs in RED_SIZES:
            if bs <= size:
                yield size, bs


def u32_input(size):
    return fmix32(np.arange(size, dtype=np.uint32))


def u64_input(size):
    return u32_input(size).astype(np.uint64)


def f32_input(size, salt=0):
    """uniform [0, 1) with 24 random bits"""
    h = fmix32(np.arange(size, dtype=np.uint32) ^ np.uint32(salt))
    return ((h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))


def mask_input(size, density, salt=0x9E3779B9):
    h = fmix32(np.arange(size, dtype=np.uint32) ^ np.uint32(salt))
    thr = np.uint64(min(int(density * 2 ** 32), 2 ** 32))
    return (h.astype(np.uint64) < thr).astype(np.uint8)


def key_input(size, buckets, skew=False):
    h = fmix32(np.arange(size, dtype=np.uint32))
    keys = (h % np.uint32(buckets)).astype(np.uint32)
    if skew:  # vcall-like: 90 % of the lanes call instance 1
        h2 = fmix32(np.arange(size, dtype=np.uint32) >> np.uint32(1))
        hot = (h % np.uint32(100)) < 90
        keys = np.where(hot, np.uint32(min(1, buckets - 1)),
                        (h2 % np.uint32(buckets))).astype(np.uint32)
    return keys


# tests/reductions.cpp:270-273 and :316-320: sizes 23*i^3+1
def cubic_sizes(n):
    return [23 * i * i * i + 1 for i in range(n)]
