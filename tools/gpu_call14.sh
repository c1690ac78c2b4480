#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scatter.py -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/test14.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test14.log)"
tail -15 gpurun_out/test14.log
timeout 300 python tools/perf_probe.py scatter
