#!/bin/bash
# round 2, late: 8- and 4-GPU bench lines of cfg3 with the final code (NCCL reduce, the default) + per-split breakdown at 8 ranks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { # name, nproc, steps, env...
  local name=$1 np=$2 steps=$3; shift 3
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $np --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2s8b_$name.json 2> gpurun_out/r2s8b_$name.err
  grep -h "split timing" gpurun_out/r2s8b_$name.err | tail -1 | cut -c1-400
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s8b_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check')}, 'k4_share', d['roofline']['k4_share_of_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
except Exception as e: print('$name failed', e)
PY
}
run n8 8 10 B200GBM_FUSED_REDUCE=0
run n8_split_timing 8 3 B200GBM_FUSED_REDUCE=0 B200GBM_SPLIT_TIMING=1
run n4 4 10 B200GBM_FUSED_REDUCE=0
