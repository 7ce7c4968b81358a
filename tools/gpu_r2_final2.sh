#!/bin/bash
# round 2 final, 2 GPUs: the whole -m gpu suite (nothing skipped but the 4-rank cases) + the N=2 bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/r2y_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_pytest_2gpu.log
tail -6 gpurun_out/r2y_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2y_bench_cfg3_n2.json 2> gpurun_out/r2y_bench_cfg3_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2y_bench_cfg3_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check','bins_sample_check','hist_conservation_check')}, d['e2e']['value'])
PY
