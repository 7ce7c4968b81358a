#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_estimators.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --config cfg5 --steps 3 --warmup 2 --no-cpu-baseline --ingest device > gpurun_out/r2j_bench_cfg5.json 2> gpurun_out/r2j_bench_cfg5.err; grep "split timing" gpurun_out/r2j_bench_cfg5.err | cut -c1-400; cut -c1-200 gpurun_out/r2j_bench_cfg5.json; python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench_cfg5.json').read().strip().splitlines()[-1]); print(d.get('parity_check'), d.get('bins_sample_check'))"
