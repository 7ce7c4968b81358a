#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_smoke_final.log | cut -c1-300
