#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_scan_wide$" -s 40 -c 2 -o gpurun_out/r2k_scan_wide python bench.py --config cfg5 --rows 2000000 --steps 1 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log | cut -c1-200
