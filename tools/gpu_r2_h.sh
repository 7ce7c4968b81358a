#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_scan$" -s 20 -c 2 -o gpurun_out/r2h_scan python bench.py --config cfg2 --rows 5000000 --steps 1 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2h_ncu_scan.log 2>&1
tail -2 gpurun_out/r2h_ncu_scan.log
