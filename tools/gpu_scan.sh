#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_baseline_sizes.py -m gpu -q --timeout 300 -p no:cacheprovider -k "scan or prefix" > gpurun_out/test_scan.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_scan.log)" | tee gpurun_out/summary.txt
grep -E "FAILED|Error|error|assert" gpurun_out/test_scan.log | head -20
timeout 200 python tools/perf_probe.py scan scan_seg 2>&1 | grep prefix | tee gpurun_out/perf_probe_scan.log
