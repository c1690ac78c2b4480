#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/test_sharded_late.log 2>&1
echo "sharded tests rc=$? $(tail -1 gpurun_out/test_sharded_late.log)"
grep -E "FAILED|Error|error|assert" gpurun_out/test_sharded_late.log | head -20
