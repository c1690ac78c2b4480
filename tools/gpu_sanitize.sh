#!/bin/bash
# compute-sanitizer over the kernels touched in round 2 (memcheck on selected parity tests, racecheck on the mkperm
# ranking kernel and the tier-2 client)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$@" -m gpu -x -q --timeout 550 -p no:cacheprovider > gpurun_out/memcheck_$name.log 2>&1; echo "memcheck $name rc=$? $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/memcheck_$name.log | tail -2 | tr '\n' ' ')"; }
run reduce tests/test_gpu_reduce.py -k "pow2 or misaligned or u8 or entry_point"
run scan tests/test_gpu_scan.py -k "seeded or carry or inplace"
run mkperm tests/test_gpu_compress_mkperm.py -k "blocked or wide or skewed or call_reduce"
run scatter tests/test_gpu_scatter.py -k "float or f16 or inc"
run packet tests/test_gpu_scatter.py -k "packet_scatter_and_gather or index_types or packet_f16"
cat > /tmp/rc_mkperm.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import drjit_core_b200 as dr
dr.jit_init()
for n, B in ((40000, 16), (100000, 1000), (70000, 70000)):
    k = torch.randint(0, B, (n,), device="cuda", dtype=torch.int32)
    perm = torch.empty(n, device="cuda", dtype=torch.int32)
    offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
    uq = dr.jit_block_mkperm(1, k, n, n, B, perm, offs)
    ks = k[perm.long()]
    assert bool((ks[1:] >= ks[:-1]).all()), "not sorted"
    x = torch.rand(n, device="cuda"); out = torch.empty(n, device="cuda")
    dr.jit_reduce(1, 14, 1, x, n, out)
    dr.jit_block_prefix_reduce(1, 14, 1, n, 256, 1, 0, x, out)
torch.cuda.synchronize()
print("racecheck workload ok")
PY
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/rc_mkperm.py > gpurun_out/racecheck_mkperm.log 2>&1; echo "racecheck mkperm/reduce/scan rc=$? $(grep -E 'workload ok|RACECHECK SUMMARY' gpurun_out/racecheck_mkperm.log | tail -2 | tr '\n' ' ')"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 ./tests/cpp/jit_h_client > gpurun_out/racecheck_client.log 2>&1; echo "racecheck client rc=$? $(grep -E 'checks|RACECHECK SUMMARY' gpurun_out/racecheck_client.log | tail -2 | tr '\n' ' ')"
