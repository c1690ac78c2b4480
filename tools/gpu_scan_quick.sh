#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scan.py -m gpu -q --timeout 120 -p no:cacheprovider -k "two_stream or full_size or carry" > gpurun_out/test_scan.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_scan.log)" | tee gpurun_out/summary.txt
grep -E "FAILED|Error|error|assert" gpurun_out/test_scan.log | head -20
