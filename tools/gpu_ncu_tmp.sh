#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'scan_ahead|scan_stream|reduce_chunks' -o gpurun_out/prof_r2_scan python tools/ncu_targets.py scan reduce > gpurun_out/ncu_r2_scan.log 2>&1; echo "ncu rc=$?"
