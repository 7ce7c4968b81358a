#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_estimators.py -m gpu -x -q 2>&1 | tail -6
