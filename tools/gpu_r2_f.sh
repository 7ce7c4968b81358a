#!/bin/bash
# round 2, call F (2 GPUs): full suite incl. multi-rank tests (new: p2p all-reduce, wide numerical, lambdarank v2) + cfg4/cfg5 lines + N=2 reduce-mode A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -14 gpurun_out/r2f_pytest.log
timeout 400 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err; tail -2 gpurun_out/r2f_bench_cfg4.err; cut -c1-330 gpurun_out/r2f_bench_cfg4.json
for mode in 0 2; do
  B200GBM_FUSED_REDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$mode bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_n2_mode$mode.json 2> gpurun_out/r2f_bench_n2_mode$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_n2_mode$mode.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','histogram_reduce','parity_check')})
PY
done
