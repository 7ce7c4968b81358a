#!/bin/bash
# column-major copy for the partition kernel: parity tests, then cfg3 / cfg2 with and without it
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ingest_scale.py tests/test_gpu_wide.py -m gpu -q -x -k "not offsets_beyond" 2>&1 | tail -3
for cc in 1 0; do
  B200GBM_COLUMN_COPY=$cc B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2o_cfg3_cc$cc.json 2> gpurun_out/r2o_cfg3_cc$cc.err
  echo "column_copy=$cc"; grep "split timing" gpurun_out/r2o_cfg3_cc$cc.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/r2o_cfg3_cc$cc.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['partition_column_copy_gb'], d['device_free_gb'], d['timed_model'])"
done
for cc in 1 0; do
  B200GBM_COLUMN_COPY=$cc timeout 600 python bench.py --config cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2o_cfg2_cc$cc.json 2> gpurun_out/r2o_cfg2_cc$cc.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2o_cfg2_cc$cc.json').read().strip().splitlines()[-1]); print('cfg2 cc=$cc', d['value'], d['ms_per_step'], d['partition_column_copy_gb'], d['timed_model'])"
done
