#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6
NG=$(nvidia-smi -L | wc -l)
echo "=== bench 100M x 512 on $NG GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 10 --warmup 3 2>gpurun_out/bn.err | grep -E '^\{' > gpurun_out/bn.json
python - <<PY
import json
d=json.load(open("gpurun_out/bn.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, {k:d["roofline"][k] for k in ("cells_per_s","frac","k4_share_of_step")}, d["e2e"]["value"], d["histogram_reduce"])
PY
tail -3 gpurun_out/bn.err
