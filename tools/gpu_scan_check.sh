#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scan.py -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/test_scan.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test_scan.log)"
timeout 300 python tools/perf_probe.py scan | grep prefix
