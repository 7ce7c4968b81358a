#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export NCCL_DEBUG=WARN
for fused in 1 0; do
echo "=== bench 100M x 512 on $NG GPUs fused=$fused"
B200GBM_FUSED_REDUCE=$fused timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2953$fused bench.py --gpus $NG --steps 10 --warmup 3 --ingest device 2>gpurun_out/s8_$fused.err | grep -E '^\{' > gpurun_out/s8_$fused.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8_$fused.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus","histogram_reduce")}, "k4 share", d["roofline"]["k4_share_of_step"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s8_$fused.err").read()[-1500:])
PY
done
