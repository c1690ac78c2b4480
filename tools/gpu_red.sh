#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_vs_reference_cuda.py -m gpu -x -q -k "reduce" --timeout 600 -p no:cacheprovider > gpurun_out/test_red.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test_red.log)"
tail -5 gpurun_out/test_red.log
python tools/perf_probe.py reduce | grep block_reduce
