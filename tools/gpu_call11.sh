#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -x -q -k mkperm --timeout 600 -p no:cacheprovider > gpurun_out/test11.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test11.log)"
timeout 300 python tools/perf_probe.py mkperm
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/launches11.csv python tools/ncu_targets.py mkperm > gpurun_out/ncu11.log 2>&1
