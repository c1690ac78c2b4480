#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"compress_pack|compress_expand|mkperm_rank_place|mkperm_tile_hist|scatter_reduce" -c 24 -o gpurun_out/prof_r1_final -f python tools/ncu_targets.py compress mkperm scatter > gpurun_out/ncu18.log 2>&1; echo "ncu rc=$?"
