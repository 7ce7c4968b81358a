#!/bin/bash
# ncu --set full of k_partition at the cfg3 shape (root and the next splits of one tree)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_partition -s 31 -c 5 -o gpurun_out/r2p_partition python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log | cut -c1-200
