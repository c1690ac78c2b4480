#!/bin/bash
# development iteration: mkperm + scatter/packet parity tests, mkperm probe at both tile sizes, ncu of the ranking kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -q --timeout 600 -p no:cacheprovider -k "mkperm or call_reduce" > gpurun_out/test_mk.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_mk.log)" | tee gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_vs_reference_cuda.py tests/test_gpu_baseline_sizes.py -m gpu -q --timeout 600 -p no:cacheprovider -k "mkperm" > gpurun_out/test_mk2.log 2>&1
echo "tests2 rc=$? $(tail -1 gpurun_out/test_mk2.log)" | tee -a gpurun_out/summary.txt
timeout 300 python tools/perf_probe.py mkperm > gpurun_out/perf_probe_mk.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/perf_probe_mk.log
for d in 7 8; do
  echo "== dbg=$d"
  B200_MKPERM_DBG=$d timeout 120 python tools/perf_probe.py mkperm 2>&1 | grep -E "mkperm B=16 \(no"
done 2>&1 | tee gpurun_out/mk_dbg.log
B200_MKPERM_TILE=4096 timeout 300 python tools/perf_probe.py mkperm > gpurun_out/perf_probe_mk32.log 2>&1; echo "probe4096 rc=$?" | tee -a gpurun_out/summary.txt
grep "mkperm B" gpurun_out/perf_probe_mk32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mkperm_rank_place|mkperm_tile_hist' -o gpurun_out/prof_mk python tools/ncu_targets.py mkperm > gpurun_out/ncu_mk.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
grep -E "FAILED|Error|error" gpurun_out/test_mk.log gpurun_out/test_mk2.log | head -20
