#!/bin/bash
mkdir -p gpurun_out
B200_MKPERM_WIDE=2 timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "mkperm or call_reduce" > gpurun_out/test_mk_wide2.log 2>&1
echo "tests wide=2 rc=$? $(tail -1 gpurun_out/test_mk_wide2.log)" | tee gpurun_out/summary.txt
for w in ${MK_MODES:-1 2}; do
  echo "== B200_MKPERM_WIDE=$w"
  B200_MKPERM_WIDE=$w timeout 300 python tools/perf_probe.py mkperm 2>&1 | grep -E "mkperm B" 
done | tee gpurun_out/perf_probe_mk.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mkperm_rank_wide' -o gpurun_out/prof_mkw python tools/ncu_targets.py mkperm256 > gpurun_out/ncu_mkw.log 2>&1; echo "ncu rc=$?"
