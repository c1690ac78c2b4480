#!/bin/bash
mkdir -p gpurun_out
for t in 0 1 3 7 11 0 3; do
  echo "--- tune=$t"
  B200_REDUCE_TUNE=$t timeout 120 python tools/perf_probe.py reduce 2>&1 | grep -E "block_reduce" | grep -E "bs=1024|bs=65536|bs=1048576|bs=268435456|bs=3145728|bs=100000 "
done | tee gpurun_out/perf_probe_red2.log
