#!/bin/bash
# very last evidence run of round 2 (2 GPUs): default bench line at N=1 and at N=2 with the final code
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f3_cfg3_n1.json 2> gpurun_out/r2f3_cfg3_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f3_cfg3_n2.json 2> gpurun_out/r2f3_cfg3_n2.err
python - <<PY
import json
for n in ('n1','n2'):
    d=json.loads(open('gpurun_out/r2f3_cfg3_%s.json'%n).read().strip().splitlines()[-1])
    print(n, {k:d.get(k) for k in ('value','ms_per_step','parity_check','bins_sample_check','hist_conservation_check','partition_column_copy_gb')}, 'frac', d['roofline']['frac'], 'k4_share', d['roofline']['k4_share_of_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['samples'])
PY
