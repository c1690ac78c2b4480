#!/bin/bash
mkdir -p gpurun_out
export B200_LIB_PATH=$PWD/gpurun_tuning.so
timeout 600 python tools/tune_scan.py 0:0 19:0 20:0 21:0 0:2 19:2 20:2 2>&1 | tee gpurun_out/tune_scan.log
