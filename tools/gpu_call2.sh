#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_vs_reference_cuda.py -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_vs_ref_cuda.log 2>&1
echo "vs_ref_cuda rc=$? $(tail -1 gpurun_out/test_vs_ref_cuda.log)" | tee gpurun_out/summary2.txt
timeout 600 python tools/ref_cuda_bench.py > gpurun_out/ref_cuda_bench.log 2>&1; echo "refbench rc=$?" | tee -a gpurun_out/summary2.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_probe.csv python tools/perf_probe.py compress mkperm scatter > gpurun_out/ncu_probe.log 2>&1; echo "ncu probe rc=$?" | tee -a gpurun_out/summary2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?" | tee -a gpurun_out/summary2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 30 -c 2 -o gpurun_out/prof_scan_stream -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/ref_cuda_bench.log
tail -30 gpurun_out/test_vs_ref_cuda.log
