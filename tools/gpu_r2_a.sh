#!/bin/bash
# round 2, call A: full 1-GPU test suite + default bench line + the other configs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; tail -3 gpurun_out/r2a_bench_cfg3.err; cat gpurun_out/r2a_bench_cfg3.json
timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; tail -3 gpurun_out/r2a_bench_cfg2.err; cat gpurun_out/r2a_bench_cfg2.json
timeout 400 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err; tail -3 gpurun_out/r2a_bench_cfg4.err; cat gpurun_out/r2a_bench_cfg4.json
