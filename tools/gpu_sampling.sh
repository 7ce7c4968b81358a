#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimators.py -m gpu -x -q 2>&1 | tail -12
timeout 500 python tools/perf_sanity.py > gpurun_out/perf_sanity2.txt 2>&1
grep -E '^\{' gpurun_out/perf_sanity2.txt | cut -c1-260
B200GBM_SPLIT_TIMING=1 timeout 500 python tools/perf_sanity.py 2>&1 | grep "split timing" | cut -c1-600
