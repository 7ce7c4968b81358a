#!/bin/bash
# round 2, call B: full 1-GPU suite on the fused partition kernel + small-shape per-split timing + bench lines
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
B200GBM_SPLIT_TIMING=1 timeout 300 python tools/perf_sanity.py > gpurun_out/r2b_perf_sanity.txt 2>&1; cat gpurun_out/r2b_perf_sanity.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err; tail -3 gpurun_out/r2b_bench_cfg3.err; cat gpurun_out/r2b_bench_cfg3.json
timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2b_bench_cfg2.json 2> gpurun_out/r2b_bench_cfg2.err; tail -3 gpurun_out/r2b_bench_cfg2.err; cat gpurun_out/r2b_bench_cfg2.json
