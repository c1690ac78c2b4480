#!/bin/bash
# Round-1 re-entry: full GPU validation + A/B of the two compress paths.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
B200_COMPRESS_PATH=2 timeout 600 python tools/perf_probe.py compress > gpurun_out/perf_probe_twopass.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
grep -E "compress|mkperm|scatter|histogram" gpurun_out/perf_probe.log
cat gpurun_out/perf_probe_twopass.log
tail -5 gpurun_out/test_gpu.log
