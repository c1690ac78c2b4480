#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py -m gpu -q --timeout 600 -p no:cacheprovider -k "mkperm or call_reduce" 2>&1 | tail -3
B200_MKPERM_PROF=1 B200_MKPERM_PROF_DUMP=1 timeout 120 python tools/perf_probe.py mkperm 2>&1 | grep -E "rk prof|mkperm B" | tail -30 | tee gpurun_out/mk_prof.log
