#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"compress_stream" -c 2 -o gpurun_out/prof_cs2 -f python tools/ncu_targets.py compress > gpurun_out/ncu9.log 2>&1; echo "ncu rc=$?"
