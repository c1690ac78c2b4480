#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress_mkperm.py tests/test_gpu_reference_suite.py -m gpu -x -q -k "mkperm or vcall or reductions" --timeout 800 -p no:cacheprovider > gpurun_out/test_mk.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test_mk.log)"
tail -5 gpurun_out/test_mk.log
python tools/mkperm_groups_probe.py
python tools/perf_probe.py mkperm | grep mkperm
