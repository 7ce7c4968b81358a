#!/bin/bash
# round 2, call D: K4 direct-load microbenchmark vs production kernel; ncu --set full of k_scan / k_partition; cfg4 launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in "ubench_hist 1" "ubench_hist_dl16 2" "ubench_hist_dl20 2" "ubench_hist_dl24 2"; do
  set -- $v
  echo "== $1 variant $2" >> gpurun_out/r2d_ubench.txt
  timeout 300 ./build/$1 10000000 $2 1 >> gpurun_out/r2d_ubench.txt 2>&1
done
cat gpurun_out/r2d_ubench.txt | grep -v "^ \"t\|smem_scatter\|^ }," 
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_partition" -s 20 -c 4 -o gpurun_out/r2d_scan_part python bench.py --config cfg2 --rows 5000000 --steps 1 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2d_ncu_scan.log 2>&1
tail -2 gpurun_out/r2d_ncu_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 500 --csv --log-file gpurun_out/r2d_launches_cfg4_5M.csv python bench.py --config cfg4 --rows 5000000 --steps 2 --warmup 1 --no-verify --no-cpu-baseline --ingest device > gpurun_out/r2d_ncu_cfg4.log 2>&1
tail -2 gpurun_out/r2d_ncu_cfg4.log
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_metrics.py tests/test_gpu_parity.py -m gpu -q -k "wide or metric or too_many or lambdarank or categorical" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --config cfg5 --steps 5 --warmup 2 > gpurun_out/r2d_bench_cfg5.json 2> gpurun_out/r2d_bench_cfg5.err; tail -5 gpurun_out/r2d_bench_cfg5.err; cat gpurun_out/r2d_bench_cfg5.json
