#!/bin/bash
# development iteration: compress parity tests, density sweep, ncu of the two compress kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_baseline_sizes.py -m gpu -q --timeout 600 -p no:cacheprovider -k "compress" > gpurun_out/test_cp.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_cp.log)" | tee gpurun_out/summary.txt
timeout 300 python tools/perf_probe.py compress_sweep > gpurun_out/perf_probe_cp.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/perf_probe_cp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'compress_pack|compress_expand' -c 6 -o gpurun_out/prof_cp python tools/ncu_targets.py compress > gpurun_out/ncu_cp.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
grep -E "FAILED|Error|error" gpurun_out/test_cp.log | head -20
