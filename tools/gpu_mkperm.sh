#!/bin/bash
# mkperm iteration: parity tests (default digit plan and wide digits forced), timing probe per plan
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "mkperm or call_reduce" > gpurun_out/test_mk.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_mk.log)" | tee gpurun_out/summary.txt
B200_MKPERM_WIDE=2 timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "mkperm or call_reduce" > gpurun_out/test_mk_wide2.log 2>&1
echo "tests wide=2 rc=$? $(tail -1 gpurun_out/test_mk_wide2.log)" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_vs_reference_cuda.py tests/test_gpu_baseline_sizes.py -m gpu -q --timeout 600 -p no:cacheprovider -k "mkperm" > gpurun_out/test_mk2.log 2>&1
echo "tests2 rc=$? $(tail -1 gpurun_out/test_mk2.log)" | tee -a gpurun_out/summary.txt
for w in 0 1 2; do
  echo "== B200_MKPERM_WIDE=$w"
  B200_MKPERM_WIDE=$w timeout 300 python tools/perf_probe.py mkperm 2>&1 | grep -E "mkperm B" 
done | tee gpurun_out/perf_probe_mk.log
if [ -n "$MK_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mkperm_rank|mkperm_tile_hist' -o gpurun_out/prof_mk python tools/ncu_targets.py mkperm > gpurun_out/ncu_mk.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
fi
grep -E "FAILED|Error|error" gpurun_out/test_mk.log gpurun_out/test_mk2.log gpurun_out/test_mk_wide2.log | head -20
