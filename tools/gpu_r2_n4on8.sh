#!/bin/bash
# 4 ranks on the 8-GPU box: is the slow non-K4 part of the earlier N=4 line NCCL's (4 of 8 GPUs) or the box's?
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { # name, nproc, steps, env...
  local name=$1 np=$2 steps=$3; shift 3
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $np --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2n48_$name.json 2> gpurun_out/r2n48_$name.err
  grep -h "split timing" gpurun_out/r2n48_$name.err | tail -1 | cut -c1-400
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n48_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check')}, 'k4_share', d['roofline']['k4_share_of_step'], 'e2e', d['e2e']['value'], d['clocks'])
except Exception as e: print('$name failed', e)
PY
}
run n4_nccl_timing 4 6 B200GBM_FUSED_REDUCE=0 B200GBM_SPLIT_TIMING=1
run n4_p2p_timing 4 6 B200GBM_FUSED_REDUCE=2 B200GBM_SPLIT_TIMING=1
