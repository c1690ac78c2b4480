#!/bin/bash
mkdir -p gpurun_out
python tools/cs_trace.py 0.01 > gpurun_out/cs_trace_001.txt 2>&1
grep -A14 "resolve:" gpurun_out/cs_trace_001.txt
tail -14 gpurun_out/cs_trace_001.txt
