#!/bin/bash
mkdir -p gpurun_out
for d in 0 8; do
  echo "--- B200_SCAN_DBG=$d"
  B200_SCAN_DBG=$d timeout 300 python tools/perf_probe.py scan 2>&1 | grep prefix | grep -E "bs=26|bs=n" 
done | tee gpurun_out/perf_probe_scan.log
export B200_LIB_PATH=$PWD/gpurun_tuning.so
timeout 600 python tools/tune_scan.py 0:2 0:10 2>&1 | tee gpurun_out/tune_scan.log
