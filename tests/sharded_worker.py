"""Worker of tests/test_gpu_sharded.py (one process per GPU under torchrun): the
NCCL / peer-mailbox sharded reduce, whole-array prefix reduction and mkperm histogram
on real GPUs against the CPU oracle on the same global input.

u32 / u64 / i32: bit-exact.  f32 Add: every element within 1e-5 (relative, floor 1)
of an fp64 accumulation.  Exit code 0 = all checks passed on this rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import oracle  # noqa: E402  (the checker)
from cases import f32_input, key_input, u32_input, u64_input  # noqa: E402
from util import to_dev  # noqa: E402

VT, OP = oracle.VT, oracle.OP


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    import drjit_core_b200 as dr
    from importlib import import_module
    sharded = import_module("drjit_core_b200.sharded")
    dr.jit_init()
    O = oracle.Oracle()
    bad = []

    for exchange in ("peer", "nccl"):
        sh = sharded.Sharded(device=dev, exchange=exchange)
        assert (sh.peer is not None) == (exchange == "peer")
        # totals: one small (carry path of the nccl front end), one large (tile-seeded)
        for total in (100003, (1 << 21) + 12345):
            start, n_local = sharded.shard_bounds(total, world, rank)
            for tname, gen in (("u32", u32_input), ("u64", u64_input), ("f32", f32_input),
                               ("i32", lambda n: u32_input(n).view(np.int32))):
                vt = VT[tname]
                dt = oracle.NP_OF_VT[vt]
                x = gen(total)
                d_x = to_dev(x[start:start + n_local])
                ops = ("add", "min", "max") if tname in ("u32", "i32") else ("add",)
                for opn in ops:
                    op = OP[opn]
                    # ---- reduce: every rank gets the global result
                    d_o = torch.zeros(4, dtype=torch.int64, device=dev)
                    sh.reduce(vt, op, d_x, n_local, d_o)
                    torch.cuda.synchronize()
                    got = d_o.cpu().numpy().view(dt)[0]
                    if tname == "f32":
                        ref = float(x.astype(np.float64).sum())
                        if abs(float(got) - ref) > 1e-5 * abs(ref):
                            bad.append((exchange, "reduce", tname, opn, total, float(got), ref))
                    else:
                        ref = O.block_reduce(vt, op, x, total)[0]
                        if got != ref:
                            bad.append((exchange, "reduce", tname, opn, total, int(got), int(ref)))
                    # ---- dot product of two equally sharded arrays (float types)
                    if tname == "f32" and opn == "add":
                        y = x[::-1].copy()
                        d_y = to_dev(y[start:start + n_local])
                        d_o = torch.zeros(4, dtype=torch.int64, device=dev)
                        sh.reduce_dot(vt, d_x, d_y, n_local, d_o)
                        torch.cuda.synchronize()
                        got = float(d_o.cpu().numpy().view(dt)[0])
                        ref = float(np.dot(x.astype(np.float64), y.astype(np.float64)))
                        if abs(got - ref) > 1e-5 * abs(ref):
                            bad.append((exchange, "dot", tname, total, got, ref))
                    # ---- whole-array prefix reduction: fwd / rev x exclusive / inclusive
                    for excl in (1, 0):
                        for rev in (0, 1):
                            d_out = torch.empty_like(d_x)
                            sh.prefix_reduce(vt, op, d_x, n_local, bool(excl), bool(rev), d_out)
                            torch.cuda.synchronize()
                            got = d_out.cpu().numpy().view(dt)
                            if tname == "f32":
                                x64 = x.astype(np.float64)
                                inc = np.cumsum(x64[::-1])[::-1] if rev else np.cumsum(x64)
                                ref = (inc - (x64 if excl else 0))[start:start + n_local]
                                err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) if n_local else 0
                                if err > 1e-5:
                                    bad.append((exchange, "scan", tname, opn, total, excl, rev, float(err)))
                            else:
                                ref = O.block_prefix_reduce(vt, op, x, total, excl, rev)[start:start + n_local]
                                if not np.array_equal(got, ref):
                                    bad.append((exchange, "scan", tname, opn, total, excl, rev))
        # ---- block-cyclic layout (peer exchange only): global block b lives on rank b % world;
        # one chained pass, block totals through the peer-mapped tables
        if exchange == "peer":
            for tname, gen, block in (("u32", u32_input, 8192), ("f32", f32_input, 16384), ("u64", u64_input, 4096),
                                      ("u32", u32_input, 1 << 18)):
                vt = VT[tname]
                dt = oracle.NP_OF_VT[vt]
                for rounds in (1, 5, 16):
                    n_local = rounds * block
                    total = n_local * world
                    x = gen(total)
                    mine = np.concatenate([x[(j * world + rank) * block:(j * world + rank + 1) * block]
                                           for j in range(rounds)])
                    d_x = to_dev(mine)
                    for excl in (1, 0):
                        d_out = torch.empty_like(d_x)
                        sh.prefix_reduce_cyclic(vt, OP["add"], d_x, n_local, block, bool(excl), d_out)
                        torch.cuda.synchronize()
                        got = d_out.cpu().numpy().view(dt)
                        if tname == "f32":
                            x64 = x.astype(np.float64)
                            full = np.cumsum(x64) - (x64 if excl else 0)
                        else:
                            full = O.block_prefix_reduce(vt, OP["add"], x, total, excl, 0)
                        ref = np.concatenate([full[(j * world + rank) * block:(j * world + rank + 1) * block]
                                              for j in range(rounds)])
                        if tname == "f32":
                            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))
                            if err > 1e-5:  # fp32 Add: 1e-5 relative against fp64
                                bad.append(("cyclic", tname, block, rounds, excl, float(err)))
                        elif not np.array_equal(got, ref):
                            bad.append(("cyclic", tname, block, rounds, excl))
        # ---- mkperm histogram (+ this rank's offsets inside every bucket)
        for total, buckets in ((100003, 16), ((1 << 21) + 5, 1024), ((1 << 21) + 5, 65536)):
            start, n_local = sharded.shard_bounds(total, world, rank)
            k = key_input(total, buckets, skew=(buckets == 1024))
            d_k = to_dev(k[start:start + n_local])
            glob, before = sh.mkperm_histogram(d_k, n_local, buckets, want_offsets=True)
            glob2 = sh.mkperm_histogram(d_k, n_local, buckets)
            torch.cuda.synchronize()
            ref = np.bincount(k, minlength=buckets).astype(np.int64)
            ref_before = np.bincount(k[:start], minlength=buckets).astype(np.int64)
            if not np.array_equal(glob.cpu().numpy().astype(np.int64), ref) or \
                    not np.array_equal(glob2.cpu().numpy().astype(np.int64), ref) or \
                    not np.array_equal(before.cpu().numpy().astype(np.int64), ref_before):
                bad.append((exchange, "histogram", total, buckets))
        sh.close()

    flag = torch.tensor([len(bad)], dtype=torch.int32, device=dev)
    dist.all_reduce(flag)
    if bad:
        print(f"rank {rank}: FAILED {bad[:10]}", flush=True)
    elif rank == 0:
        print(f"SHARDED_OK world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
