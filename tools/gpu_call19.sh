#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jit_h_client.py tests/test_gpu_reduce.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test19.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test19.log)"
time ./tests/cpp/jit_h_client
compute-sanitizer --tool memcheck ./tests/cpp/jit_h_client 2>&1 | tail -5
