#!/bin/bash
# Full GPU validation pass: smoke, every GPU test, per-primitive probe, bench, ncu launch list + DRAM traffic of the bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/perf_probe.log
tail -5 gpurun_out/test_gpu.log
