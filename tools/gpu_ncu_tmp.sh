#!/bin/bash
mkdir -p gpurun_out
B200_SCAN_LAG=2048 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_ahead' -o gpurun_out/prof_ahead python tools/ncu_targets.py scan_whole > gpurun_out/ncu_ahead.log 2>&1; echo "ncu rc=$?"
