#!/bin/bash
# GPU tests + probe only (no bench)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
grep -E "FAILED|Error|passed|failed" gpurun_out/test_gpu.log | tail -30
