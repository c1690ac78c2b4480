#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_vs_reference_cuda.py -m gpu -q -k "scatter" --timeout 900 -p no:cacheprovider > gpurun_out/test_vsref.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test_vsref.log)"
tail -25 gpurun_out/test_vsref.log
