#!/bin/bash
# round 2, call C: full 1-GPU suite (wide features, device metrics, lambdarank kernel) + cfg4/cfg5 bench lines + launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -40 gpurun_out/r2c_pytest.log
timeout 400 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_cfg4.json 2> gpurun_out/r2c_bench_cfg4.err; tail -3 gpurun_out/r2c_bench_cfg4.err; cat gpurun_out/r2c_bench_cfg4.json
timeout 900 python bench.py --config cfg5 --steps 5 --warmup 2 > gpurun_out/r2c_bench_cfg5.json 2> gpurun_out/r2c_bench_cfg5.err; tail -5 gpurun_out/r2c_bench_cfg5.err; cat gpurun_out/r2c_bench_cfg5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r2c_launches_cfg2_5M.csv python bench.py --config cfg2 --rows 5000000 --steps 2 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2c_ncu_cfg2.log 2>&1
tail -2 gpurun_out/r2c_ncu_cfg2.log
