#!/bin/bash
# round 2, call E: full 1-GPU suite after the scan / partition rework + bench lines + launch list + sanitizer
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -12 gpurun_out/r2e_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r2e_launches_cfg2_5M.csv python bench.py --config cfg2 --rows 5000000 --steps 2 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2e_ncu_cfg2.log 2>&1
timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2e_bench_cfg2.json 2> gpurun_out/r2e_bench_cfg2.err; tail -3 gpurun_out/r2e_bench_cfg2.err; cat gpurun_out/r2e_bench_cfg2.json | cut -c1-400
timeout 400 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2e_bench_cfg4.json 2> gpurun_out/r2e_bench_cfg4.err; tail -3 gpurun_out/r2e_bench_cfg4.err; cat gpurun_out/r2e_bench_cfg4.json | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_cfg3.json 2> gpurun_out/r2e_bench_cfg3.err; tail -3 gpurun_out/r2e_bench_cfg3.err; cat gpurun_out/r2e_bench_cfg3.json | cut -c1-400
bash tools/gpu_r2_sanitize.sh
