#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mkperm_rank_wide' -o gpurun_out/prof_mkw python tools/ncu_targets.py mkperm256 > gpurun_out/ncu_mkw.log 2>&1; echo "ncu rc=$?"
