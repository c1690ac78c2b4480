#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
echo "== single launch"; timeout 300 python tools/perf_probe.py reduce 2>&1 | grep -E "bs=1048576|bs=268435456|reduce u32|bs=65536"
echo "== two launches"; B200_REDUCE_TWO_LAUNCHES=1 timeout 300 python tools/perf_probe.py reduce 2>&1 | grep -E "bs=1048576|bs=268435456|reduce u32|bs=65536"
done
timeout 600 python -m pytest tests/test_gpu_reduce.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
