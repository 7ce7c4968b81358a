#!/bin/bash
# one gpurun call: parity tests + small bench + full bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
echo "=== bench small (2M x 512)"
timeout 600 python bench.py --rows 2000000 --steps 5 --warmup 2 --cpu-sample-rows 500000 2>&1 | tail -5
