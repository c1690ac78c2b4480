"""Generate tests/golden/reference_cpu.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

Every entry is produced by the reference's own CPU implementation of the path
(oracle/_ref/libref_cpu.so: LLVMThreadState::block_reduce / block_prefix_reduce /
compress / block_mkperm, /root/reference/src/llvm_ts.cpp:265-933) on the
deterministic inputs of tests/golden/cases.py.  Outputs are stored as SHA-1
digests plus a few leading values (the arrays themselves would be hundreds of
megabytes).  Floating-point results are NOT pinned here: the reference has no
floating-point reduction test (tests/reductions.cpp:408-414 is #if 0) and its
CPU sums depend on the thread count; those are checked by tolerance instead.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from cases import (RED_SIZES, cubic_sizes, key_input, mask_input, red_pairs,  # noqa: E402
                   u32_input, u64_input)

VT, OP = oracle.VT, oracle.OP


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def entry(a):
    return {"sha1": digest(a), "n": int(a.shape[0]),
            "head": [int(x) for x in a[:4]]}


def main():
    R = oracle.Reference()
    out = {"_meta": {"source": "mitsuba-renderer/drjit-core @ 9d9d6e2, CPU (LLVM-backend) "
                               "primitives via oracle/_ref/libref_cpu.so",
                     "threads": R.threads, "red_sizes": RED_SIZES}}

    # tests/reductions.cpp:109-267 -- block sum / prefix sum over the size grid
    for size, bs in red_pairs():
        for tname, x in (("u32", u32_input(size)), ("u64", u64_input(size))):
            vt = VT[tname]
            out[f"block_reduce/{tname}/add/{size}/{bs}"] = entry(
                R.block_reduce(vt, OP["add"], x, bs))
            for excl in (0, 1):
                for rev in (0, 1):
                    out[f"prefix/{tname}/add/{size}/{bs}/{excl}/{rev}"] = entry(
                        R.block_prefix_reduce(vt, OP["add"], x, bs, excl, rev))

    # other operators / signed types on a smaller grid (not covered by the
    # reference's tests; pinned against its CPU implementation)
    for size, bs in ((1000, 7), (1000, 1000), (163880, 333), (163880, 163880), (4097, 64)):
        x32 = u32_input(size)
        for tname, x in (("u32", x32), ("i32", x32.view(np.int32)),
                         ("u64", (x32.astype(np.uint64) << np.uint64(17)) ^ x32.astype(np.uint64)),
                         ("i64", ((x32.astype(np.uint64) << np.uint64(33)) ^ x32.astype(np.uint64)).view(np.int64))):
            for opn in ("add", "mul", "min", "max", "and_", "or_"):
                out[f"ops_reduce/{tname}/{opn}/{size}/{bs}"] = entry(
                    R.block_reduce(VT[tname], OP[opn], x, bs))
                out[f"ops_prefix/{tname}/{opn}/{size}/{bs}/1/0"] = entry(
                    R.block_prefix_reduce(VT[tname], OP[opn], x, bs, 1, 0))
                out[f"ops_prefix/{tname}/{opn}/{size}/{bs}/0/1"] = entry(
                    R.block_prefix_reduce(VT[tname], OP[opn], x, bs, 0, 1))

    # tests/reductions.cpp:269-313 -- compress
    for size in cubic_sizes(30):
        for dens in (0.0, 0.01, 0.5, 0.99, 1.0):
            m = mask_input(size, dens)
            idx, cnt = R.compress(m)
            e = entry(idx)
            e["count"] = int(cnt)
            out[f"compress/{size}/{dens}"] = e

    # tests/reductions.cpp:315-406 -- mkperm (block_size == size)
    for size in cubic_sizes(30)[::3]:
        for buckets in (1, 2, 16, 24, 1024, 5000, 65536):
            k = key_input(size, buckets)
            perm, offs, uq = R.block_mkperm(k, size, buckets)
            e = entry(perm)
            e["unique"] = int(uq)
            e["offsets_sha1"] = digest(offs[:4 * uq])
            out[f"mkperm/{size}/{buckets}"] = e
    # blocked variant (newer API; groups that divide evenly or are small: the
    # reference's CPU task split reads out of bounds for a ragged last group
    # larger than 16384 elements, src/llvm_ts.cpp:846-850)
    for size, bs, buckets in ((100000, 1000, 16), (100000, 12500, 300), (65536, 4096, 7),
                              (99999, 333, 40)):
        k = key_input(size, buckets)
        perm, _, uq = R.block_mkperm(k, bs, buckets)
        e = entry(perm)
        e["unique"] = int(uq)
        out[f"mkperm_blocked/{size}/{bs}/{buckets}"] = e

    # all / any (tests/reductions.cpp:78-107)
    for size in (1, 3, 4, 5, 1000, 4099):
        f = np.zeros(size, dtype=np.uint8)
        t = np.ones(size, dtype=np.uint8)
        out[f"allany/{size}"] = {"all_t": R.all(t), "any_t": R.any(t),
                                 "all_f": R.all(f), "any_f": R.any(f)}
        f[size // 2] = 1
        t[size // 2] = 0
        out[f"allany_flip/{size}"] = {"all_t": R.all(t), "any_t": R.any(t),
                                      "all_f": R.all(f), "any_f": R.any(f)}

    # identities (src/var.cpp:2642-2652)
    for vt in (4, 7, 8, 9, 10, 13, 14, 15):
        for op in range(1, 7):
            out[f"identity/{vt}/{op}"] = int(R.reduce_identity(vt, op))

    path = os.path.join(HERE, "reference_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f"wrote {len(out)} entries to {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
