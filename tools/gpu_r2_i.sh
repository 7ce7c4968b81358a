#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --config cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2i_bench_cfg5.json 2> gpurun_out/r2i_bench_cfg5.err; grep "split timing" gpurun_out/r2i_bench_cfg5.err | cut -c1-400
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2i_bench_cfg2.json 2> gpurun_out/r2i_bench_cfg2.err; grep "split timing" gpurun_out/r2i_bench_cfg2.err | cut -c1-400
timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2i_bench_cfg2_plain.json 2>/dev/null; cut -c1-200 gpurun_out/r2i_bench_cfg2_plain.json
timeout 600 python -m pytest tests/test_gpu_estimators.py -m gpu -q -x -k "multiclassova" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r2i_launches_cfg5.csv python bench.py --config cfg5 --rows 2000000 --steps 1 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2i_ncu_cfg5.log 2>&1; tail -1 gpurun_out/r2i_ncu_cfg5.log | cut -c1-100
