"""Quick per-primitive timing at the BASELINE.json sizes (development aid; the
numbers that count come from bench.py)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr  # noqa: E402

CUDA = 1
F32, U32 = 14, 8
ADD = 1
PEAK = 6450.6


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    want = set(sys.argv[1:])

    def on(name):
        return not want or name in want

    dr.jit_init()
    res = {}
    n = 1 << 28
    x = torch.rand(n, device="cuda", dtype=torch.float32)
    out = torch.empty(n, device="cuda", dtype=torch.float32)

    def report(name, ms, nbytes):
        res[name] = {"ms_med": ms[0], "ms_min": ms[1], "GBs": nbytes / ms[0] / 1e6,
                     "frac": nbytes / ms[0] / 1e6 / PEAK}
        print(f"{name:40s} {ms[0]:8.3f} ms  {nbytes / ms[0] / 1e6:8.1f} GB/s  {nbytes / ms[0] / 1e6 / PEAK:5.2f}", flush=True)

    ms = timeit(lambda: out.copy_(x))
    report("torch copy (r+w)", ms, 8 * n)
    for bs in [1, 2, 4, 16, 128, 256, 1024, 4096, 1 << 16, 1 << 20, n, 3, 7, 100, 1000, 100000, 3 << 20] if on("reduce") else []:
        ms = timeit(lambda: dr.jit_block_reduce(CUDA, F32, ADD, n, bs, x, out))
        report(f"block_reduce f32 bs={bs}", ms, 4 * n * (1 + 1 / bs))
    for bs in [1, 2, 16, 128, 256, 1024, 4096, 8192, 1 << 16, n, 3, 100, 1000, 100000, 3 << 20] if on("scan") else []:
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, x, out))
        report(f"prefix f32 excl bs={bs}", ms, 8 * n if bs > 1 else 4 * n)
    for bs in [1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 20, 1 << 24] if "scan_seg" in want else []:
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, x, out))
        report(f"prefix f32 excl bs={bs}", ms, 8 * n)
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 0, 1, x, out))
        report(f"prefix f32 incl reverse bs={bs}", ms, 8 * n)
    xi = x.view(torch.int32)
    oi = out.view(torch.int32)
    if on("scan"):
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, U32, ADD, n, n, 1, 0, xi, oi))
        report("prefix u32 excl bs=n", ms, 8 * n)
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, U32, ADD, n, n, 0, 1, xi, oi))
        report("prefix u32 incl reverse bs=n", ms, 8 * n)
        x64 = x.view(torch.int64)
        o64 = out.view(torch.int64)
        ms = timeit(lambda: dr.jit_block_prefix_reduce(CUDA, 10, ADD, n // 2, n // 2, 1, 0, x64, o64))
        report("prefix u64 excl bs=n (2^27)", ms, 8 * n)
    if on("reduce"):
        ms = timeit(lambda: dr.jit_reduce(CUDA, U32, ADD, xi, n, oi))
        report("reduce u32", ms, 4 * n)
        ms = timeit(lambda: dr.jit_reduce_dot(CUDA, F32, x, out, n, oi))
        report("dot f32", ms, 8 * n)

    dens = (0.001, 0.01, 0.03, 0.05, 0.1, 0.5, 0.99) if "compress_sweep" in want else (0.01, 0.5, 0.99)
    for d in dens if on("compress") or "compress_sweep" in want else []:
        m = (torch.rand(n, device="cuda") < d).to(torch.uint8)
        cnt = int(m.sum().item())
        ms = timeit(lambda: dr.jit_compress(CUDA, m, n, oi))
        report(f"compress d={d}", ms, n + 4 * cnt)
        cdev = torch.zeros(1, device="cuda", dtype=torch.int32)
        ms = timeit(lambda: dr.compress_async(m, n, oi, cdev))
        report(f"compress_async d={d}", ms, n + 4 * cnt)
        del m

    n2 = 1 << 26
    perm = torch.empty(n2, device="cuda", dtype=torch.int32)
    for B in (16, 256, 1024, 65536) if on("mkperm") else []:
        k = torch.randint(0, B, (n2,), device="cuda", dtype=torch.int32)
        offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
        ms = timeit(lambda: dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, offs), iters=5)
        report(f"mkperm B={B}", ms, 8 * n2)
        ms = timeit(lambda: dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, None), iters=5)
        report(f"mkperm B={B} (no offsets)", ms, 8 * n2)
        h = torch.empty(B, device="cuda", dtype=torch.int32)
        ms = timeit(lambda: dr.mkperm_histogram(k, n2, B, h), iters=5)
        report(f"histogram B={B}", ms, 4 * n2)
    if on("mkperm"):
        # skewed callee ids: one dominant bucket (what vectorised dispatch often sees)
        B = 16
        offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
        k = torch.zeros(n2, device="cuda", dtype=torch.int32)
        ms = timeit(lambda: dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, offs), iters=5)
        report("mkperm B=16 all keys equal", ms, 8 * n2)
        k = torch.where(torch.rand(n2, device="cuda") < 0.9, 3, torch.randint(0, B, (n2,), device="cuda")).to(torch.int32)
        ms = timeit(lambda: dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, offs), iters=5)
        report("mkperm B=16 90% one bucket", ms, 8 * n2)
        del k

    if not on("scatter"):
        return
    m2 = 1 << 20
    idx = torch.randint(0, m2, (n2,), device="cuda", dtype=torch.int32)
    val = torch.rand(n2, device="cuda")
    tgt = torch.zeros(m2, device="cuda")
    for mode, name in ((1, "direct"), (2, "local")):
        ms = timeit(lambda: dr.scatter_reduce(F32, ADD, tgt, val, idx, None, n2, mode=mode), iters=5)
        report(f"scatter_add f32 random {name}", ms, 8 * n2)
    idx2 = (torch.arange(n2, device="cuda", dtype=torch.int32) >> 6) & (m2 - 1)
    for mode, name in ((1, "direct"), (2, "local")):
        ms = timeit(lambda: dr.scatter_reduce(F32, ADD, tgt, val, idx2, None, n2, mode=mode), iters=5)
        report(f"scatter_add f32 coherent {name}", ms, 8 * n2)

    ms = timeit(lambda: dr.scatter_reduce(F32, ADD, tgt, val, idx2, None, n2, mode=0), iters=5)
    report("scatter_add f32 coherent auto", ms, 8 * n2)
    ms = timeit(lambda: dr.scatter_reduce(F32, ADD, tgt, val, idx, None, n2, mode=0), iters=5)
    report("scatter_add f32 random auto", ms, 8 * n2)
    # scatter_inc: one shared counter (queue compaction) and 2^20 random counters
    cnt = torch.zeros(m2, device="cuda", dtype=torch.int32)
    old = torch.empty(n2, device="cuda", dtype=torch.int32)
    zero = torch.zeros(n2, device="cuda", dtype=torch.int32)
    ms = timeit(lambda: dr.scatter_inc(cnt, zero, None, old, n2), iters=5)
    report("scatter_inc one counter", ms, 8 * n2)
    ms = timeit(lambda: dr.scatter_inc(cnt, idx, None, old, n2), iters=5)
    report("scatter_inc random counters", ms, 8 * n2)
    # packet scatter: 2^24 RGBA splats into 2^20 x 4 floats
    n3 = 1 << 24
    comps = [torch.rand(n3, device="cuda") for _ in range(4)]
    tgt4 = torch.zeros(4 * m2, device="cuda")
    for mode, name in ((1, "direct"), (0, "auto")):
        ms = timeit(lambda: dr.scatter_reduce_packet(F32, ADD, tgt4, comps, idx, None, n3, mode=mode), iters=5)
        report(f"scatter_add_packet f32x4 random {name} (2^24)", ms, 20 * n3)
    def four_scalar():
        for k in range(4):
            dr.scatter_reduce(F32, ADD, tgt4, comps[k], idx, None, n3, mode=1)
    ms = timeit(four_scalar, iters=5)
    report("  same as 4 scalar scatters (layout differs)", ms, 32 * n3)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
