#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== ncu full K4 v3 (10M x 512)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_hist_build_ws -c 3 -o gpurun_out/k4v3_prof python bench.py --rows 10000000 --steps 1 --warmup 1 --no-cpu-baseline --ingest device > gpurun_out/ncu_k4v3_bench.log 2>&1; tail -1 gpurun_out/ncu_k4v3_bench.log | cut -c1-200
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v3.csv python bench.py --rows 10000000 --steps 2 --warmup 1 --no-cpu-baseline --ingest device > gpurun_out/ncu_launch_v3.log 2>&1
echo "=== bench 10M x 256 regression-like? (cfg2 shape, binary objective)"
timeout 300 python bench.py --rows 10000000 --features 256 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
