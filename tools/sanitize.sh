#!/bin/bash
# tools/sanitize.sh TAG — on the GPU box: memcheck and racecheck over every kernel family (tools/sanitize.py)
T=${1:-rXX}
for tool in memcheck racecheck; do
  BACON_IVP_GRID=2 timeout 400 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/${T}_sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all families ran|Error|AssertionError" gpurun_out/${T}_sanitize_$tool.log | head -5
done
