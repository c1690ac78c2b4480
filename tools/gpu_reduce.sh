#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_vs_reference_cuda.py -m gpu -q --timeout 900 -p no:cacheprovider -k "reduce" > gpurun_out/test_red.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_red.log)" | tee gpurun_out/summary.txt
grep -E "FAILED|^E  " gpurun_out/test_red.log | head -20
timeout 300 python tools/perf_probe.py reduce 2>&1 | grep -E "block_reduce" | tee gpurun_out/perf_probe_red.log
