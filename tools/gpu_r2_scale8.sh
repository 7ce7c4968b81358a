#!/bin/bash
# round 2: 8-GPU scaling of cfg3 (NCCL vs fused reduce), per-split breakdown at 8 ranks, and the 4-rank parity tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2s8_gpus.txt
run() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/r2s8_$name.json 2> gpurun_out/r2s8_$name.err
  tail -n 3 gpurun_out/r2s8_$name.err | grep -i "split timing\|error" | cut -c1-600
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s8_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check')}, 'k4_share', d['roofline']['k4_share_of_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
except Exception as e: print('$name failed', e)
PY
}
run nccl B200GBM_FUSED_REDUCE=0
run p2p B200GBM_FUSED_REDUCE=2
run fused B200GBM_FUSED_REDUCE=1
STEPS=3 run nccl_split_timing B200GBM_FUSED_REDUCE=0 B200GBM_SPLIT_TIMING=1
grep -h "split timing" gpurun_out/r2s8_nccl_split_timing.err | head -2 | cut -c1-500
STEPS=3 run p2p_split_timing B200GBM_FUSED_REDUCE=2 B200GBM_SPLIT_TIMING=1
grep -h "split timing" gpurun_out/r2s8_p2p_split_timing.err | head -2 | cut -c1-500
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "4" > gpurun_out/r2s8_pytest_4rank.log 2>&1; tail -4 gpurun_out/r2s8_pytest_4rank.log
