#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_estimators.py -m gpu -x -q 2>&1 | tail -15
NG=$(nvidia-smi -L | wc -l)
for fused in 1 0; do
echo "=== bench 40M x 512 on $NG GPUs fused=$fused"
B200GBM_FUSED_REDUCE=$fused timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2952$fused bench.py --gpus $NG --rows 40000000 --steps 8 --warmup 2 --ingest device 2>gpurun_out/m2_$fused.err | grep -E '^\{' > gpurun_out/m2_$fused.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m2_$fused.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "k4 share", d["roofline"]["k4_share_of_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/m2_$fused.err").read()[-1500:])
PY
done
