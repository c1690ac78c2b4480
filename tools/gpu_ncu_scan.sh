#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan_stream|reduce_rows|reduce_chunks" -s 20 -c 8 -o gpurun_out/prof_scan_final -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_scan.log 2>&1; echo "ncu rc=$?"
