"""compute-sanitizer workload for the kernels of late round 2 (development aid): tile-group scans, the two-stream
chained scan (in place, reverse, ragged), reduce with small odd / many large blocks.  Results are checked against torch."""
import sys

import torch

sys.path.insert(0, ".")
import drjit_core_b200 as dr  # noqa: E402

CUDA, I32, U32, F32, ADD = 1, 7, 8, 14, 1
dr.jit_init()
g = torch.Generator(device="cuda").manual_seed(1)


def scan_ref(x, bs, excl, rev):
    n = x.numel()
    pad = (-n) % bs
    if rev:
        x = x.flip(0)
        # blocks are anchored at the physical start: in reversed order the ragged block comes first
        x = torch.cat([torch.zeros(pad, dtype=x.dtype, device=x.device), x])
    else:
        x = torch.cat([x, torch.zeros(pad, dtype=x.dtype, device=x.device)])
    c = x.view(-1, bs).cumsum(1, dtype=x.dtype)
    if excl:
        c = c - x.view(-1, bs)
    c = c.reshape(-1)
    c = c[pad:] if rev else c[:n]
    return c.flip(0) if rev else c


tile = 8192
for n, bs, excl, rev, inplace in (
        (40 * tile + 5, 2 * tile, 1, 0, False), (64 * tile, 16 * tile, 0, 1, False),       # tile groups
        (1100 * tile + 77, 1100 * tile + 77, 1, 0, False), (1100 * tile + 77, 1100 * tile + 77, 0, 1, True),
        (1100 * tile + 77, 32 * tile, 1, 1, False), (1030 * tile, 64 * tile, 0, 0, True)):  # two streams
    x = torch.randint(-1000, 1000, (n,), device="cuda", dtype=torch.int32, generator=g)
    ref = scan_ref(x, bs, excl, rev)
    out = x.clone() if inplace else torch.empty_like(x)
    dr.jit_block_prefix_reduce(CUDA, I32, ADD, n, bs, excl, rev, out if inplace else x, out)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), (n, bs, excl, rev, inplace)
for n, bs in ((1 << 22, 3), (1 << 22, 7), (1000003, 100), (1 << 24, 1 << 18), (3 << 22, 1 << 20), (1 << 24, 1 << 24)):
    x = torch.randint(-1000, 1000, (n,), device="cuda", dtype=torch.int32, generator=g)
    nb = (n + bs - 1) // bs
    out = torch.empty(nb, device="cuda", dtype=torch.int32)
    dr.jit_block_reduce(CUDA, I32, ADD, n, bs, x, out)
    pad = (-n) % bs
    ref = torch.cat([x, torch.zeros(pad, dtype=x.dtype, device="cuda")]).view(-1, bs).sum(1, dtype=torch.int32)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), (n, bs)
print("late round-2 workload ok")
