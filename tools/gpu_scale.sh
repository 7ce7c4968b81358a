#!/bin/bash
# strong-scaling run of the headline workload on all GPUs of the box + parity on 4 ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export NCCL_DEBUG=WARN
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for n in $NG; do
  echo "=== bench 100M x 512 on $n GPUs"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_$n.err | grep -E '^\{' > gpurun_out/scale_$n.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$n.json"))
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "k4 frac", d["roofline"]["frac"], "share", d["roofline"]["k4_share_of_step"], "e2e", d["e2e"]["value"], "ingest_ms", d["e2e"]["ingest_ms"], "clocks", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
