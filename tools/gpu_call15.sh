#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -x -q -k mkperm --timeout 600 -p no:cacheprovider > gpurun_out/test15.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test15.log)"
timeout 900 python tools/ref_cuda_bench.py > gpurun_out/ref_cuda_bench.log 2>&1; echo "refbench rc=$?"
cat gpurun_out/ref_cuda_bench.log | tail -30
