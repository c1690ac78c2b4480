#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -q --timeout 900 -p no:cacheprovider -k "odd_blocks" > gpurun_out/test_scan_odd.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/test_scan_odd.log)"
grep -E "^E  " gpurun_out/test_scan_odd.log | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_stream_kernel' -o gpurun_out/prof_sg python tools/ncu_targets.py scan_general > gpurun_out/ncu_sg.log 2>&1; echo "ncu rc=$?"
