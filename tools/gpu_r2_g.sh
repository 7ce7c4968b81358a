#!/bin/bash
# round 2, call G: per-split breakdown at the N=8 shard size on one GPU; quick regression of the partition change
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --rows 12500000 --steps 5 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2g_bench_12p5M.json 2> gpurun_out/r2g_bench_12p5M.err; grep "split timing" gpurun_out/r2g_bench_12p5M.err | cut -c1-400; cut -c1-300 gpurun_out/r2g_bench_12p5M.json
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2g_bench_cfg2.json 2> gpurun_out/r2g_bench_cfg2.err; grep "split timing" gpurun_out/r2g_bench_cfg2.err | cut -c1-400; cut -c1-300 gpurun_out/r2g_bench_cfg2.json
timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2g_bench_cfg2_plain.json 2>/dev/null; cut -c1-200 gpurun_out/r2g_bench_cfg2_plain.json
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py tests/test_gpu_ingest_scale.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; tail -4 gpurun_out/r2g_pytest.log
