#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tune_scan.py 3 8 9 10 11 12 > gpurun_out/tune_scan.log 2>&1
cat gpurun_out/tune_scan.log | grep -E "^geom|CTAs/SM"
