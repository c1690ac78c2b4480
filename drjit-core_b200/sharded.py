"""Multi-GPU front end: large reductions, whole-array prefix scans and the
mkperm histogram sharded across the GPUs of one box (one process per GPU,
torch.distributed / NCCL over NVLink for the small exchanges).

The reference has no multi-GPU code on this path (no NCCL/MPI anywhere; one
CUDADevice + stream per GPU, src/cuda_core.cpp:352-516) -- this is the new
functionality BASELINE.json's north_star asks for.  Data layout: rank r owns
the contiguous shard [r * n_local, (r + 1) * n_local) of the global array;
compress and scatter-reduce stay per device.

Exchange steps (payloads are a few bytes to a few KiB, i.e. latency bound):
  reduce     local reduce -> all_gather of the W partials -> every rank
             combines them in rank order (bit-exact for integers, and the same
             fixed order on every rank for floating point).
  dot        the same with a local dot product in front (rank-order sum).
  scan       local reduce -> all_gather of the W totals -> exclusive scan of
             the totals (tiny, on device) -> local scan seeded with the rank's
             carry.  12 B/element per GPU instead of 8, so the ceiling against a
             1-GPU single-pass scan is about W * 8 / 12.  For large shards the
             reduce pass leaves one sum per 32 KiB tile (block_reduce), whose
             exclusive scan (+ carry) gives every tile its prefix: the scan pass
             then runs WITHOUT a look-back chain (b200_prefix_reduce_seeded,
             6.5 instead of 5.5 TB/s).  Small or misaligned shards use the
             chained single pass with a carry (b200_prefix_reduce_carry).
  histogram  local per-bucket counts -> all_reduce(sum); optional exclusive
             scan over ranks gives each rank its global output offsets.

`local_ops` is the object that executes the per-GPU primitives; it defaults to
the CUDA library.  (tests/ inject a stand-in to exercise the exchange logic
under gloo on CPU-only machines.)
"""
import torch
import torch.distributed as dist

from . import (JitBackend, ReduceOp, TYPE_SIZE, VarType, jit_block_prefix_reduce,
               jit_block_reduce, jit_reduce, jit_reduce_dot, mkperm_histogram, prefix_reduce_carry,
               prefix_reduce_seeded, scan_tile_elems)


def value_size(vt):
    """Bytes of the arithmetic ("value") type: float16 is carried as float32."""
    return 4 if vt == VarType.Float16 else TYPE_SIZE[vt]


class CudaLocalOps:
    """Per-GPU primitives from libdrjit_core_b200.so on the current stream."""

    def reduce(self, vt, op, in_, size, out):
        jit_reduce(JitBackend.CUDA, vt, op, in_, size, out)

    def reduce_dot(self, vt, a, b, size, out):
        jit_reduce_dot(JitBackend.CUDA, vt, a, b, size, out)

    def block_reduce(self, vt, op, size, block_size, in_, out):
        jit_block_reduce(JitBackend.CUDA, vt, op, size, block_size, in_, out)

    def block_prefix_reduce(self, vt, op, size, block_size, exclusive, reverse, in_, out):
        jit_block_prefix_reduce(JitBackend.CUDA, vt, op, size, block_size, exclusive,
                                reverse, in_, out)

    def prefix_reduce_carry(self, vt, op, size, exclusive, reverse, in_, out, carry_in,
                            carry_out):
        prefix_reduce_carry(vt, op, size, exclusive, reverse, in_, out, carry_in, carry_out)

    def histogram(self, values, size, bucket_count, hist):
        mkperm_histogram(values, size, bucket_count, hist)

    def scan_tile_elems(self, vt, in_, out):
        """Tile size of the seeded scan, or 0 when it cannot serve these arrays."""
        if vt == VarType.Float16 or in_.data_ptr() % 16 or out.data_ptr() % 16:
            return 0
        return scan_tile_elems(vt)

    def prefix_reduce_seeded(self, vt, op, size, exclusive, reverse, in_, out, seeds):
        prefix_reduce_seeded(vt, op, size, exclusive, reverse, in_, out, seeds)


def shard_bounds(total, world, rank):
    """Contiguous, balanced partition of `total` elements (first ranks get the
    remainder), as (start, length)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class Sharded:
    """exchange="peer": the collectives are ONE C-ABI call per rank (b200_sharded_*):
    totals travel through peer-mapped mailboxes over NVLink, no NCCL call and no
    allocation on the data path (csrc/sharded.cu).  exchange="nccl": the local passes
    are C-ABI calls, the totals are gathered with torch.distributed (also what the
    gloo tests on CPU exercise with a stand-in for the local passes).  "auto": peer
    mailboxes on CUDA when the IPC mapping succeeds on every rank, else nccl."""

    def __init__(self, group=None, device=None, local_ops=None, seeded_min_tiles=64,
                 exchange="auto"):
        self.seeded_min_tiles = seeded_min_tiles
        self.peer = None
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
            else torch.device("cpu"))
        self.ops = local_ops if local_ops is not None else CudaLocalOps()
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"exchange must be auto, peer or nccl (got {exchange!r})")
        if exchange != "nccl" and local_ops is None and self.device.type == "cuda":
            self._connect_peers(required=(exchange == "peer"))

    def _connect_peers(self, required):
        from . import ShardedContext

        def gather(handle):
            if self.world == 1:
                return [handle]
            out = [None] * self.world
            dist.all_gather_object(out, handle, group=self.group)
            return out

        err = None
        try:
            self.peer = ShardedContext(self.rank, self.world, gather)
        except (RuntimeError, AssertionError) as e:
            err = e
        # every rank must take the same path
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if self.peer is not None:
                self.peer.close()
                self.peer = None
            if required:
                raise RuntimeError(f"sharded: peer mailboxes could not be mapped on every rank ({err})")

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def _bytes(self, n):
        return torch.zeros(n, dtype=torch.uint8, device=self.device)

    def _gather(self, mine):
        """all_gather of one small byte tensor per rank -> (world * len) bytes"""
        out = self._bytes(mine.numel() * self.world)
        if self.world == 1:
            out.copy_(mine)
        else:
            dist.all_gather_into_tensor(out, mine, group=self.group)
        return out

    # -- reduce -------------------------------------------------------------
    def reduce(self, vt, op, local_in, local_size, out):
        """Whole-array reduction of the global array; every rank receives the
        result in `out` (device scalar of type vt)."""
        if self.peer is not None:
            self.peer.reduce(vt, op, local_in, local_size, out)
            return
        tsize = TYPE_SIZE[vt]
        partial = self._bytes(tsize)
        if local_size > 0:
            self.ops.reduce(vt, op, local_in, local_size, partial)
        else:
            self._fill_identity(partial, vt, op)
        gathered = self._gather(partial)
        self.ops.block_reduce(vt, op, self.world, self.world, gathered, out)

    def reduce_dot(self, vt, local_a, local_b, local_size, out):
        """Dot product of two equally sharded float arrays; every rank receives the
        result in `out` (device scalar of type vt; rank-order sum of the partials)."""
        if self.peer is not None:
            self.peer.reduce_dot(vt, local_a, local_b, local_size, out)
            return
        partial = self._bytes(TYPE_SIZE[vt])
        if local_size > 0:
            self.ops.reduce_dot(vt, local_a, local_b, local_size, partial)
        else:
            partial.zero_()
        gathered = self._gather(partial)
        self.ops.block_reduce(vt, ReduceOp.Add, self.world, self.world, gathered, out)

    def _fill_identity(self, buf, vt, op):
        from . import jit_reduce_identity
        ident = jit_reduce_identity(vt, op)
        raw = ident.to_bytes(8, "little")[:buf.numel()]
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))

    # -- scan -----------------------------------------------------------------
    def prefix_reduce(self, vt, op, local_in, local_size, exclusive, reverse, local_out):
        """Whole-array prefix reduction of the global array (block_size ==
        global size); rank r's shard of the result is written to local_out."""
        if vt == VarType.Float16:
            raise RuntimeError("sharded prefix_reduce(): float16 is not supported "
                               "(per-shard totals would be rounded to half)")
        if (self.peer is not None and local_in.data_ptr() % 16 == 0 and local_out.data_ptr() % 16 == 0):
            # (alignment is a property of the allocation: the same on every rank)
            self.peer.prefix_reduce(vt, op, local_size, exclusive, reverse, local_in, local_out)
            return
        tsize = TYPE_SIZE[vt]
        total = self._bytes(tsize)
        # large shard: the reduce pass also leaves one sum per scan tile
        tile = self.ops.scan_tile_elems(vt, local_in, local_out) if local_size > 0 else 0
        ntiles = -(-local_size // tile) if tile else 0
        seeded = ntiles >= max(1, self.seeded_min_tiles)
        if seeded:
            tsums = self._bytes(ntiles * tsize)
            self.ops.block_reduce(vt, op, local_size, min(tile, local_size), local_in, tsums)
            self.ops.reduce(vt, op, tsums, ntiles, total)
        elif local_size > 0:
            self.ops.reduce(vt, op, local_in, local_size, total)
        else:
            self._fill_identity(total, vt, op)
        totals = self._gather(total)
        # exclusive scan over ranks; a reverse scan accumulates from the last rank
        carries = self._bytes(tsize * self.world)
        self.ops.block_prefix_reduce(vt, op, self.world, self.world, True, bool(reverse),
                                     totals, carries)
        carry = carries[self.rank * tsize:(self.rank + 1) * tsize]
        if seeded:
            # exclusive prefix of every tile = rank carry + tiles in front of it
            seeds = self._bytes(ntiles * tsize)
            self.ops.prefix_reduce_carry(vt, op, ntiles, True, reverse, tsums, seeds, carry, None)
            self.ops.prefix_reduce_seeded(vt, op, local_size, exclusive, reverse, local_in,
                                          local_out, seeds)
        elif local_size > 0:
            self.ops.prefix_reduce_carry(vt, op, local_size, exclusive, reverse, local_in,
                                         local_out, carry, None)

    # -- mkperm histogram -------------------------------------------------------
    def prefix_reduce_cyclic(self, vt, op, local_in, local_size, block_size, exclusive, local_out):
        """Whole-array prefix reduction over a BLOCK-CYCLIC layout (global block b of
        `block_size` elements lives on rank b % world as local block b // world): one pass
        over the data instead of the two that contiguous shards need.  Peer mailboxes only
        (b200_sharded_prefix_reduce_cyclic, csrc/sharded.cu)."""
        if self.peer is None:
            raise RuntimeError("sharded prefix_reduce_cyclic(): needs the peer-mailbox exchange "
                               "(CUDA, mailboxes mapped on every rank)")
        self.peer.prefix_reduce_cyclic(vt, op, local_size, block_size, exclusive, local_in, local_out)

    def mkperm_histogram(self, local_values, local_size, bucket_count, want_offsets=False):
        """Global per-bucket counts (int32 tensor of bucket_count entries on every
        rank).  With want_offsets also returns this rank's exclusive offset per
        bucket among the ranks (counts of lower ranks), which together with the
        exclusive scan of the global counts gives its global output slots."""
        if self.peer is not None and bucket_count <= 65536:
            glob = torch.empty(bucket_count, dtype=torch.int32, device=self.device)
            before = torch.empty(bucket_count, dtype=torch.int32, device=self.device) if want_offsets else None
            self.peer.histogram(local_values, local_size, bucket_count, glob, before)
            return (glob, before) if want_offsets else glob
        local = torch.zeros(bucket_count, dtype=torch.int32, device=self.device)
        if local_size > 0:
            self.ops.histogram(local_values, local_size, bucket_count, local)
        if self.world == 1:
            return (local, torch.zeros_like(local)) if want_offsets else local
        if not want_offsets:
            glob = local.clone()
            dist.all_reduce(glob, op=dist.ReduceOp.SUM, group=self.group)
            return glob
        allh = torch.zeros(self.world * bucket_count, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(allh, local, group=self.group)
        allh = allh.view(self.world, bucket_count)
        glob = allh.sum(dim=0, dtype=torch.int32)
        before = allh[:self.rank].sum(dim=0, dtype=torch.int32)
        return glob, before
