#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$@" -m gpu -x -q --timeout 550 -p no:cacheprovider > gpurun_out/memcheck_$name.log 2>&1; echo "memcheck $name rc=$? $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/memcheck_$name.log | tail -2 | tr '\n' ' ')"; }
run reduce tests/test_gpu_reduce.py -k "pow2 or misaligned or u8 or entry_point"
run scan tests/test_gpu_scan.py -k "seeded or carry or inplace"
run mkperm tests/test_gpu_compress_mkperm.py -k "blocked or wide or skewed"
run scatter tests/test_gpu_scatter.py -k "float or f16 or inc"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 ./tests/cpp/jit_h_client > gpurun_out/racecheck_client.log 2>&1; echo "racecheck client rc=$? $(grep -E 'checks|RACECHECK SUMMARY' gpurun_out/racecheck_client.log | tail -2 | tr '\n' ' ')"
