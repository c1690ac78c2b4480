#!/bin/bash
# First GPU validation: smoke, parity tests per file, perf probe.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
for f in reduce scan compress_mkperm scatter; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "test_$f rc=$? $(tail -1 gpurun_out/test_$f.log)" | tee -a gpurun_out/summary.txt
done
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
tail -60 gpurun_out/perf_probe.log
