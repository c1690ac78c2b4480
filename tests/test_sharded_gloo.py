"""CPU-only, world_size 2 over gloo: the exchange logic of the multi-GPU front
end (drjit-core_b200/sharded.py).  The per-GPU primitives are replaced by a
stand-in built on the CPU oracle -- this exercises partitioning, the gathers
and the carry computation, not the CUDA kernels (those are covered by the
`-m gpu` tests and by tests/test_gpu_sharded.py on a GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocalOps:
    """Stand-in for CudaLocalOps on CPU byte tensors (test infrastructure)."""

    def __init__(self):
        import oracle
        self.O = oracle.Oracle()
        self.np_of = oracle.NP_OF_VT

    def _view(self, t, vt, n):
        return t.numpy().view(np.uint8)[:n * np.dtype(self.np_of[vt]).itemsize].view(self.np_of[vt])

    def reduce(self, vt, op, in_, size, out):
        self._view(out, vt, 1)[:] = self.O.block_reduce(vt, op, self._view(in_, vt, size), size)

    def block_reduce(self, vt, op, size, block_size, in_, out):
        r = self.O.block_reduce(vt, op, self._view(in_, vt, size), block_size)
        self._view(out, vt, r.size)[:] = r

    def reduce_dot(self, vt, a, b, size, out):
        self._view(out, vt, 1)[:] = self.O.reduce_dot(vt, self._view(a, vt, size), self._view(b, vt, size))

    def block_prefix_reduce(self, vt, op, size, block_size, exclusive, reverse, in_, out):
        r = self.O.block_prefix_reduce(vt, op, self._view(in_, vt, size), block_size, exclusive, reverse)
        self._view(out, vt, size)[:] = r

    def prefix_reduce_carry(self, vt, op, size, exclusive, reverse, in_, out, carry_in, carry_out):
        x = self._view(in_, vt, size)
        c = self._view(carry_in, vt, 1).copy()
        # scan of [carry, x...] in the scan direction, dropping the carry slot
        ext = np.concatenate([x, c]) if reverse else np.concatenate([c, x])
        r = self.O.block_prefix_reduce(vt, op, ext, ext.size, exclusive, reverse)
        self._view(out, vt, size)[:] = r[:-1] if reverse else r[1:]

    TILE = 16  # stand-in tile size of the seeded scan

    def scan_tile_elems(self, vt, in_, out):
        return self.TILE

    def prefix_reduce_seeded(self, vt, op, size, exclusive, reverse, in_, out, seeds):
        x = self._view(in_, vt, size)
        sd = self._view(seeds, vt, -(-size // self.TILE))
        res = np.empty_like(x)
        for t in range(sd.size):
            lo, hi = t * self.TILE, min(size, (t + 1) * self.TILE)
            c = sd[t:t + 1]
            ext = np.concatenate([x[lo:hi], c]) if reverse else np.concatenate([c, x[lo:hi]])
            r = self.O.block_prefix_reduce(vt, op, ext, ext.size, exclusive, reverse)
            res[lo:hi] = r[:-1] if reverse else r[1:]
        self._view(out, vt, size)[:] = res

    def histogram(self, values, size, bucket_count, hist):
        k = values.numpy().view(np.uint32)[:size]
        hist.copy_(torch.from_numpy(np.bincount(k, minlength=bucket_count).astype(np.int32)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cases import u32_input, f32_input, key_input
        import drjit_core_b200  # noqa: F401
        from drjit_core_b200.sharded import Sharded, shard_bounds
        VT, OP = oracle.VT, oracle.OP
        O = oracle.Oracle()
        sh = Sharded(device=torch.device("cpu"), local_ops=OracleLocalOps())
        # a second front end that takes the tile-seeded scan path for every shard
        sh_seeded = Sharded(device=torch.device("cpu"), local_ops=OracleLocalOps(), seeded_min_tiles=1)
        sh.seeded_min_tiles = 1 << 30
        ok = True
        for total in (1, 2, 7, 1000, 100003):
            start, n = shard_bounds(total, world, rank)
            x = u32_input(total)
            mine = torch.from_numpy(x[start:start + n].copy().view(np.uint8))
            # reduce: every rank gets the global result, bit-exact
            for opn in ("add", "max", "and_"):
                out = torch.zeros(4, dtype=torch.uint8)
                sh.reduce(VT["u32"], OP[opn], mine, n, out)
                ok &= out.numpy().view(np.uint32)[0] == O.block_reduce(VT["u32"], OP[opn], x, total)[0]
            # scan: rank's shard of the global result
            for excl in (0, 1):
                for rev in (0, 1):
                    ref = O.block_prefix_reduce(VT["u32"], OP["add"], x, total, excl, rev)
                    for front in (sh, sh_seeded):
                        out = torch.zeros(max(n, 1) * 4, dtype=torch.uint8)
                        front.prefix_reduce(VT["u32"], OP["add"], mine, n, excl, rev, out)
                        ok &= np.array_equal(out.numpy().view(np.uint32)[:n], ref[start:start + n])
            # histogram: global counts and the rank's offsets
            k = key_input(total, 37)
            mine_k = torch.from_numpy(k[start:start + n].copy().view(np.uint8))
            glob, before = sh.mkperm_histogram(mine_k, n, 37, want_offsets=True)
            ok &= np.array_equal(glob.numpy(), np.bincount(k, minlength=37))
            ok &= np.array_equal(before.numpy(), np.bincount(k[:start], minlength=37))
            glob2 = sh.mkperm_histogram(mine_k, n, 37)
            ok &= np.array_equal(glob2.numpy(), glob.numpy())
        # floating point: same fixed combination order on every rank
        xf = f32_input(50001)
        start, n = shard_bounds(xf.size, world, rank)
        out = torch.zeros(4, dtype=torch.uint8)
        sh.reduce(VT["f32"], OP["add"], torch.from_numpy(xf[start:start + n].copy().view(np.uint8)), n, out)
        got = out.numpy().view(np.float32)[0]
        ok &= abs(got - xf.astype(np.float64).sum()) <= 1e-4 * xf.astype(np.float64).sum()
        gathered = [None] * world
        dist.all_gather_object(gathered, float(got))
        ok &= len(set(gathered)) == 1
        # dot product of two equally sharded arrays (an empty shard contributes zero)
        for total in (1, 50001):
            xa, xb = f32_input(total), f32_input(total)[::-1].copy()
            start, n = shard_bounds(total, world, rank)
            out = torch.zeros(4, dtype=torch.uint8)
            sh.reduce_dot(VT["f32"], torch.from_numpy(xa[start:start + n].copy().view(np.uint8)),
                          torch.from_numpy(xb[start:start + n].copy().view(np.uint8)), n, out)
            got = out.numpy().view(np.float32)[0]
            ref = float(np.dot(xa.astype(np.float64), xb.astype(np.float64)))
            ok &= abs(got - ref) <= 1e-5 * abs(ref)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    import drjit_core_b200  # noqa: F401
    from drjit_core_b200.sharded import shard_bounds
    for total in (0, 1, 7, 8, 1000, 2 ** 32):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(n for _, n in spans) == total
            for (s0, n0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + n0 == s1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_sharded_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}
