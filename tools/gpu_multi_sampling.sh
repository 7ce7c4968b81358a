#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "row_sampling or (matches_oracle and binary-2-0)" 2>&1 | tail -25
