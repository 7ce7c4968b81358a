#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "=== full gpu test suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== ncu K4 DRAM traffic, one iteration at 100M x 512"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k4_hist_build_ws --launch-skip 30 --launch-count 30 --csv --log-file gpurun_out/k4_traffic.csv python bench.py --rows 100000000 --steps 1 --warmup 1 --ingest device --no-cpu-baseline > gpurun_out/k4_traffic.log 2>&1
python tools/ncu_k4_traffic.py gpurun_out/k4_traffic.csv 100000000 512 gpurun_out/r01_k4_dram_traffic_100000000x512.json
cp gpurun_out/r01_k4_dram_traffic_100000000x512.json profiles/
echo "=== launch list (ncu gpu__time_duration) 10M x 512"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --rows 10000000 --steps 2 --warmup 1 --ingest device --no-cpu-baseline > gpurun_out/launches_final.log 2>&1
echo "=== bench default N=1"
timeout 900 python bench.py 2>gpurun_out/bench_final.err | grep -E '^\{' > gpurun_out/bench_final.json; cut -c1-1800 gpurun_out/bench_final.json
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | grep -E '^\{' > gpurun_out/bench_ref.json; cut -c1-600 gpurun_out/bench_ref.json
