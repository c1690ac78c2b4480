#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "seeded or carry" --timeout 600 -p no:cacheprovider > gpurun_out/test_seeded.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test_seeded.log)"
tail -15 gpurun_out/test_seeded.log
