#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_compress_mkperm.py tests/test_gpu_scatter.py -m gpu -x -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/test8.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test8.log)" | tee gpurun_out/summary8.txt
timeout 600 python tools/perf_probe.py compress mkperm scatter > gpurun_out/perf_probe8.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary8.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/launches8.csv python tools/ncu_targets.py compress mkperm scatter > gpurun_out/ncu8.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary8.txt
cat gpurun_out/perf_probe8.log
tail -30 gpurun_out/test8.log
