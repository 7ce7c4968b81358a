#!/bin/bash
# compute-sanitizer on the hand-synchronised kernels; logs -> gpurun_out/, summaries copied to profiles/
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|done" gpurun_out/r2_sanitizer_$tool.log | head -12
done
