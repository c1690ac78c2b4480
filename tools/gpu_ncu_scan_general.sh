#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_stream_kernel' -o gpurun_out/prof_sg python tools/ncu_targets.py scan_general > gpurun_out/ncu_sg.log 2>&1; echo "ncu rc=$?"
