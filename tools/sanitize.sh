#!/bin/bash
# tools/sanitize.sh — compute-sanitizer memcheck + racecheck over tools/sanitize.py (on the GPU box); logs in gpurun_out/
export BACON_IVP_GRID=6
python tools/sanitize.py 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python tools/sanitize.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool: exit $? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
