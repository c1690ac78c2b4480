#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'reduce_tiles' -o gpurun_out/prof_ro python tools/ncu_targets.py reduce_odd > gpurun_out/ncu_ro.log 2>&1; echo "ncu rc=$?"
