#!/bin/bash
mkdir -p gpurun_out
for st in 0 1000 2500 5000 10000; do
  echo "== stagger $st"
  B200_MKPERM_STAGGER=$st timeout 300 python tools/perf_probe.py mkperm 2>&1 | grep -E "mkperm B=(16|1024|65536) \(no"
done | tee gpurun_out/mk_stagger.log
