"""Times the REFERENCE'S OWN CUDA KERNELS (oracle/_ref/libref_cuda.so: the
unmodified reference with its CUDA backend, compute_75 PTX JIT-compiled by the
driver) on the BASELINE.json configurations, next to the sm_100a kernels of this
repository, on the same device buffers.  Test / measurement infrastructure only.

Writes gpurun_out/ref_cuda_bench.json and prints one line per configuration.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr  # noqa: E402
import oracle  # noqa: E402

CUDA, F32, U32, ADD = 1, 14, 8, 1


def time_events(fn, stream, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        fn()
        e.record(stream)
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def time_wall(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    dr.jit_init()
    R = oracle.ReferenceCUDA.get()
    rstream = torch.cuda.ExternalStream(R.lib.refcuda_stream())
    ostream = torch.cuda.current_stream()
    res = {}

    def row(name, ref_ms, our_ms, nbytes):
        res[name] = {"reference_cuda_ms": ref_ms, "b200_ms": our_ms, "speedup": ref_ms / our_ms,
                     "reference_cuda_GBs": nbytes / ref_ms / 1e6, "b200_GBs": nbytes / our_ms / 1e6}
        print(f"{name:44s} reference CUDA {ref_ms:9.3f} ms   ours {our_ms:8.3f} ms   x{ref_ms / our_ms:6.2f}",
              flush=True)

    n = 1 << 28
    x = torch.rand(n, device="cuda", dtype=torch.float32)
    out = torch.empty(n, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    for bs in (2, 32, 1024, 4096, n):
        r = time_events(lambda: R.block_reduce(F32, ADD, n, bs, x.data_ptr(), out.data_ptr()), rstream)
        o = time_events(lambda: dr.jit_block_reduce(CUDA, F32, ADD, n, bs, x, out), ostream)
        row(f"block_reduce f32 2^28 bs={bs}", r, o, 4 * n * (1 + 1 / bs))
    for bs in (2, 32, 1024, 4096, n):
        r = time_events(lambda: R.block_prefix_reduce(F32, ADD, n, bs, 1, 0, x.data_ptr(), out.data_ptr()), rstream)
        o = time_events(lambda: dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, x, out), ostream)
        row(f"block_prefix_reduce f32 excl 2^28 bs={bs}", r, o, 8 * n)
    xi, oi = x.view(torch.int32), out.view(torch.int32)
    r = time_events(lambda: R.block_reduce(U32, ADD, n, n, xi.data_ptr(), oi.data_ptr()), rstream)
    o = time_events(lambda: dr.jit_reduce(CUDA, U32, ADD, xi, n, oi), ostream)
    row("reduce u32 2^28", r, o, 4 * n)
    r = time_events(lambda: R.block_prefix_reduce(U32, ADD, n, n, 1, 0, xi.data_ptr(), oi.data_ptr()), rstream)
    o = time_events(lambda: dr.jit_block_prefix_reduce(CUDA, U32, ADD, n, n, 1, 0, xi, oi), ostream)
    row("prefix sum u32 excl 2^28", r, o, 8 * n)
    r = time_events(lambda: R.reduce_dot(F32, x.data_ptr(), out.data_ptr(), n, oi.data_ptr()), rstream)
    o = time_events(lambda: dr.jit_reduce_dot(CUDA, F32, x, out, n, oi), ostream)
    row("reduce_dot f32 2^28", r, o, 8 * n)

    for d in (0.01, 0.5, 0.99):
        m = (torch.rand(n, device="cuda") < d).to(torch.uint8)
        cnt = int(m.sum().item())
        r = time_wall(lambda: R.compress(m.data_ptr(), n, oi.data_ptr()))
        o = time_wall(lambda: dr.jit_compress(CUDA, m, n, oi))
        row(f"compress 2^28 d={d}", r, o, n + 4 * cnt)
        del m

    n2 = 1 << 26
    perm = torch.empty(n2, device="cuda", dtype=torch.int32)
    for B in (16, 1024, 65536):
        k = torch.randint(0, B, (n2,), device="cuda", dtype=torch.int32)
        offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
        r = time_wall(lambda: R.block_mkperm(k.data_ptr(), n2, n2, B, perm.data_ptr(), offs.data_ptr()), iters=5)
        o = time_wall(lambda: dr.jit_block_mkperm(CUDA, k, n2, n2, B, perm, offs), iters=5)
        row(f"mkperm 2^26 B={B}", r, o, 8 * n2)

    m2 = 1 << 20
    val = torch.rand(n2, device="cuda")
    tgt = torch.zeros(m2, device="cuda")
    for kind in ("random", "coherent"):
        if kind == "random":
            idx = torch.randint(0, m2, (n2,), device="cuda", dtype=torch.int32)
        else:
            idx = (torch.arange(n2, device="cuda", dtype=torch.int32) >> 6) & (m2 - 1)
        for mode, mname in ((0, "auto"), (1, "direct")):
            # the reference JIT-compiles the fused kernel on first use (cached afterwards);
            # time 5 evaluations inside one call and divide
            R.scatter_reduce(F32, ADD, mode, tgt.data_ptr(), m2, val.data_ptr(), idx.data_ptr(), None, n2, 2)
            R.sync()
            t0 = time.perf_counter()
            R.scatter_reduce(F32, ADD, mode, tgt.data_ptr(), m2, val.data_ptr(), idx.data_ptr(), None, n2, 10)
            R.sync()
            r = (time.perf_counter() - t0) * 1e3 / 10
            o = time_events(lambda: dr.scatter_reduce(F32, ADD, tgt, val, idx, None, n2, mode=mode), ostream)
            row(f"scatter_add f32 2^26->2^20 {kind} {mname}", r, o, 8 * n2)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_cuda_bench.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
