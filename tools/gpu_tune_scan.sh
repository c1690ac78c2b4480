#!/bin/bash
mkdir -p gpurun_out
export B200_LIB_PATH=$PWD/gpurun_tuning.so
timeout 600 python tools/tune_scan.py 0:2 13:2 14:2 17:2 15:2 16:2 18:2 9:2 2>&1 | tee gpurun_out/tune_scan.log
