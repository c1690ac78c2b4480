#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mkperm_rank_place|compress_stream" -c 6 -o gpurun_out/prof_rk_cs -f python tools/ncu_targets.py compress mkperm > gpurun_out/ncu7.log 2>&1; echo "ncu rc=$?"
