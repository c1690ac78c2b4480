import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, numpy as np
import drjit_core_b200 as dr
dr.jit_init()
n = 1 << 26
perm = torch.empty(n, device="cuda", dtype=torch.int32)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(iters):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
for B in (16, 1024):
    k = torch.randint(0, B, (n,), device="cuda", dtype=torch.int32)
    for bs in (n, 1 << 22, 1 << 17, 1 << 16, 1 << 13, 1000):
        ms = timeit(lambda: dr.jit_block_mkperm(1, k, n, bs, B, perm, None))
        print(f"mkperm 2^26 B={B} block_size={bs}: {ms:.3f} ms", flush=True)
