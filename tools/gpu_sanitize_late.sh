#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sanitize_round2_late.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$? $(tail -1 gpurun_out/san_plain.log)"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_round2_late.py > gpurun_out/memcheck_late.log 2>&1; echo "memcheck late rc=$? $(grep -E 'workload ok|ERROR SUMMARY' gpurun_out/memcheck_late.log | tail -2 | tr '\n' ' ')"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_round2_late.py > gpurun_out/racecheck_late.log 2>&1; echo "racecheck late rc=$? $(grep -E 'workload ok|RACECHECK SUMMARY' gpurun_out/racecheck_late.log | tail -2 | tr '\n' ' ')"
