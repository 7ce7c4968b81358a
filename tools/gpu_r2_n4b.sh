#!/bin/bash
# final 4-GPU line of cfg3 with the batched-load partition kernel + column copy
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n4b_n4.json 2> gpurun_out/r2n4b_n4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2n4b_n4.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check','partition_column_copy_gb')}, 'k4_share', d['roofline']['k4_share_of_step'], 'e2e', d['e2e']['value'], d['clocks'])
PY
