#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
export NCCL_DEBUG=WARN
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -25
NG=$(nvidia-smi -L | wc -l)
echo "=== bench strong scaling on $NG GPUs (20M x 512 quick)"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --rows 20000000 --steps 5 --warmup 2 2>&1 | grep -E '^\{|Error|error' | cut -c1-1500
