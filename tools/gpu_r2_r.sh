#!/bin/bash
# two-level chunk scan of k_partition: mid-scale parity (leaves above 4M rows take that path) + cfg3 per-split timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ingest_scale.py tests/test_gpu_parity.py -m gpu -q -x -k "mid_scale or column_copy or heavy or binary" 2>&1 | tail -3
B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ingest device > gpurun_out/r2r_cfg3.json 2> gpurun_out/r2r_cfg3.err
grep "split timing" gpurun_out/r2r_cfg3.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/r2r_cfg3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity_check'], d['bins_sample_check'], d['hist_conservation_check'], d['timed_model'])"
