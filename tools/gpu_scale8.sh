#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
NG=$(nvidia-smi -L | wc -l)
echo "=== bench 100M x 512 on $NG GPUs"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b8.err | grep -E '^\{' > gpurun_out/b8.json
python - <<PY
import json
d=json.load(open("gpurun_out/b8.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, {k:d["roofline"][k] for k in ("cells_per_s","frac","k4_share_of_step")}, d["e2e"]["value"], d["e2e"]["ingest_ms"], d["histogram_reduce"], d["clocks"])
PY
tail -2 gpurun_out/b8.err
