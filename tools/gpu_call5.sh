#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_warp_ops > gpurun_out/ubench_warp_ops.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"compress_stream|scatter_reduce" -c 5 -o gpurun_out/prof_compress_scatter -f python tools/ncu_targets.py compress scatter > gpurun_out/ncu_cs.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/ubench_warp_ops.txt
