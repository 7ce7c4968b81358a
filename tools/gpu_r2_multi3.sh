#!/bin/bash
# the multi-rank parity tests on 2 GPUs with the final code (column copy, reworked partition)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multi.py tests/test_gpu_estimators.py tests/test_gpu_wide.py -m gpu -q -rs -k "data_parallel or rank or two_ranks or two_tasks or reduce" > gpurun_out/r2m3_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m3_pytest_multi.log
tail -6 gpurun_out/r2m3_pytest_multi.log
