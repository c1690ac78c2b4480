#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scan.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "two_stream or full_size or block_groups" > gpurun_out/test_scan.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_scan.log)" | tee gpurun_out/summary.txt
grep -E "FAILED|Error|error|assert" gpurun_out/test_scan.log | head -20
timeout 120 python tools/perf_probe.py scan 2>&1 | grep prefix | grep -E "bs=26|bs=n|65536" | tee gpurun_out/perf_probe_scan.log
