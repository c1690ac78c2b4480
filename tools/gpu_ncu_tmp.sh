#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'compress|mkperm|scatter' -o gpurun_out/prof_r2_cms python tools/ncu_targets.py compress mkperm scatter > gpurun_out/ncu_r2_cms.log 2>&1; echo "ncu rc=$?"
