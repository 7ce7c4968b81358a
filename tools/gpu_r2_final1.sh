#!/bin/bash
# round 2 final, 1 GPU: smoke, full -m gpu suite, the default bench line + reference arm, the other configs, ncu DRAM traffic of K4, launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/r2z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2z_pytest.log
tail -8 gpurun_out/r2z_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_cfg3.json 2> gpurun_out/r2z_bench_cfg3.err; tail -2 gpurun_out/r2z_bench_cfg3.err; cut -c1-250 gpurun_out/r2z_bench_cfg3.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_cfg3_reference.json 2> gpurun_out/r2z_bench_cfg3_reference.err; cut -c1-250 gpurun_out/r2z_bench_cfg3_reference.json
for c in cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2z_bench_$c.json 2> gpurun_out/r2z_bench_$c.err; tail -2 gpurun_out/r2z_bench_$c.err; cut -c1-250 gpurun_out/r2z_bench_$c.json
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k4_hist_build_ws -s 90 -c 30 --csv --log-file gpurun_out/r2z_k4_traffic_cfg3.csv python bench.py --steps 2 --warmup 3 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2z_ncu_traffic.log 2>&1
python tools/ncu_k4_traffic.py gpurun_out/r2z_k4_traffic_cfg3.csv 100000000 512 gpurun_out/r02_k4_dram_traffic_100000000x512.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r2z_launches_cfg3_10M.csv python bench.py --rows 10000000 --steps 2 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2z_ncu_launches.log 2>&1
tail -1 gpurun_out/r2z_ncu_launches.log | cut -c1-200
