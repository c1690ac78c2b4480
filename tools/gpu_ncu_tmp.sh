#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'compress|mkperm|scan_stream' --csv --log-file gpurun_out/launches_cm.csv python tools/ncu_targets.py compress mkperm > gpurun_out/ncu_cm.log 2>&1; echo "ncu rc=$?"
