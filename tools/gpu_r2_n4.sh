#!/bin/bash
# why is the 4-GPU step slower than expected? per-split breakdown with NCCL and with the P2P all-reduce kernel, NCCL's algorithm choice
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n4_topo.txt 2>&1
run() { # name, nproc, steps, env...
  local name=$1 np=$2 steps=$3; shift 3
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $np --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2n4_$name.json 2> gpurun_out/r2n4_$name.err
  grep -h "split timing" gpurun_out/r2n4_$name.err | tail -1 | cut -c1-400
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n4_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check','clocks')}, 'k4_share', d['roofline']['k4_share_of_step'], 'e2e', d['e2e']['value'])
except Exception as e: print('$name failed', e)
PY
}
run n4_nccl_timing 4 5 B200GBM_FUSED_REDUCE=0 B200GBM_SPLIT_TIMING=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL
grep -h "NCCL INFO.*\(Algo\|NVLS\|Channel\|comm.*nranks\)" gpurun_out/r2n4_n4_nccl_timing.err | head -12 | cut -c1-200
run n4_p2p_timing 4 5 B200GBM_FUSED_REDUCE=2 B200GBM_SPLIT_TIMING=1
run n4_p2p 4 10 B200GBM_FUSED_REDUCE=2
run n4_nccl 4 10 B200GBM_FUSED_REDUCE=0
run n2_nccl 2 10 B200GBM_FUSED_REDUCE=0
