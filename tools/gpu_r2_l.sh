#!/bin/bash
# round 2, late: full -m gpu suite + bench lines of every config with the final wide-scan code, and the 5M x 256 regression point
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -6 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench_cfg3.json 2> gpurun_out/r2l_bench_cfg3.err; tail -2 gpurun_out/r2l_bench_cfg3.err; cut -c1-250 gpurun_out/r2l_bench_cfg3.json
for c in cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2l_bench_$c.json 2> gpurun_out/r2l_bench_$c.err; tail -2 gpurun_out/r2l_bench_$c.err; cut -c1-250 gpurun_out/r2l_bench_$c.json
done
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --config cfg2 --rows 5000000 --steps 20 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2l_bench_cfg2_5m.json 2> gpurun_out/r2l_bench_cfg2_5m.err; grep "split timing" gpurun_out/r2l_bench_cfg2_5m.err | cut -c1-300; cut -c1-250 gpurun_out/r2l_bench_cfg2_5m.json
