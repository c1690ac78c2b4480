"""One launch of every hot kernel at the BASELINE sizes, for `ncu --set full`
captures (development aid; never a source of bench numbers).

    ncu --set full --clock-control none --import-source on -k regex:<pattern> \
        -o gpurun_out/prof python tools/ncu_targets.py [names...]

names: reduce scan compress mkperm scatter (default: all)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr  # noqa: E402

CUDA, F32, U32, ADD = 1, 14, 8, 1


def main():
    want = set(sys.argv[1:]) or {"reduce", "scan", "compress", "mkperm", "scatter"}
    dr.jit_init()
    n = 1 << 28
    x = torch.rand(n, device="cuda", dtype=torch.float32)
    out = torch.empty(n, device="cuda", dtype=torch.float32)
    oi = out.view(torch.int32)
    if "reduce" in want:
        for bs in (2, 128, 1024, 4096, n):
            dr.jit_block_reduce(CUDA, F32, ADD, n, bs, x, out)
    if "scan" in want:
        for bs in (2, 128, 1024, 4096, n):
            dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, x, out)
        dr.jit_block_prefix_reduce(CUDA, U32, ADD, n, n, 1, 0, x.view(torch.int32), oi)
    if "scan_whole" in want:
        dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, n, 1, 0, x, out)
        dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, n, 1, 0, x, out)
    if "reduce_odd" in want:
        for bs in (3, 7, 100):
            dr.jit_block_reduce(CUDA, F32, ADD, n, bs, x, out)
    if "scan_general" in want:
        for bs in (3, 1000, 100000):
            dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, x, out)
    if "compress" in want:
        for d in (0.01, 0.5, 0.99):
            m = (torch.rand(n, device="cuda") < d).to(torch.uint8)
            dr.jit_compress(CUDA, m, n, oi)
            del m
    n2 = 1 << 26
    if "mkperm" in want:
        perm = torch.empty(n2, device="cuda", dtype=torch.int32)
        for B in (16, 1024, 65536):
            k = torch.randint(0, B, (n2,), device="cuda", dtype=torch.int32)
            offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
            dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, offs)
    if "mkperm256" in want:
        perm = torch.empty(n2, device="cuda", dtype=torch.int32)
        k = torch.randint(0, 256, (n2,), device="cuda", dtype=torch.int32)
        dr.jit_block_mkperm(CUDA, k, n2, n2, 256, perm, None)
    if "scatter" in want:
        m2 = 1 << 20
        idx = torch.randint(0, m2, (n2,), device="cuda", dtype=torch.int32)
        val = torch.rand(n2, device="cuda")
        tgt = torch.zeros(m2, device="cuda")
        for mode in (1, 2):
            dr.scatter_reduce(F32, ADD, tgt, val, idx, None, n2, mode=mode)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
