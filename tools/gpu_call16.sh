#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jit_h_client.py tests/test_gpu_scatter.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test16.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test16.log)"
tail -20 gpurun_out/test16.log
./oracle/_ref/jit_h_client_refhdr | tail -2
timeout 300 python tools/perf_probe.py scatter | grep scatter_add
