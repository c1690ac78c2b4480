#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_reference_suite.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/test12.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/test12.log)"
grep -E "reference tests passed|FAILED|Error|error" gpurun_out/test12.log | head -30
