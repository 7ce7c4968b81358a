#!/bin/bash
# compute-sanitizer (memcheck, synccheck) on the final round-2 kernels: reworked k_partition + column copy, k_scan_wide selection, k4_hist_wide
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2b_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|hazard|Error|done" gpurun_out/r2b_sanitizer_$tool.log | head -8
done
