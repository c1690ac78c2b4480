"""Development aid: two launches of the POW2 streaming scan (bs=128 and bs=TILE)
for an ncu comparison."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr
dr.jit_init()
n = 1 << 28
x = torch.rand(n, device="cuda", dtype=torch.float32)
out = torch.empty(n, device="cuda", dtype=torch.float32)
for bs in (128, 4096, 128, 4096):
    dr.jit_block_prefix_reduce(1, 14, 1, n, bs, 1, 0, x, out)
torch.cuda.synchronize()
