#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_warp_ops > gpurun_out/ubench_warp_ops.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py tests/test_gpu_scatter.py -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider -k "compress or scatter" > gpurun_out/test3.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test3.log)" | tee gpurun_out/summary3.txt
timeout 600 python -m pytest tests/test_gpu_vs_reference_cuda.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "compress or scatter" > gpurun_out/test3b.log 2>&1
echo "tests vs ref rc=$? $(tail -1 gpurun_out/test3b.log)" | tee -a gpurun_out/summary3.txt
timeout 600 python tools/perf_probe.py compress scatter > gpurun_out/perf_probe3.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary3.txt
B200_COMPRESS_PATH=2 timeout 600 python tools/perf_probe.py compress > gpurun_out/perf_probe3_twopass.log 2>&1
cat gpurun_out/ubench_warp_ops.txt gpurun_out/perf_probe3.log
tail -20 gpurun_out/test3.log
