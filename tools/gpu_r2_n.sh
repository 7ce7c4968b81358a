#!/bin/bash
# A/B of the partition kernel's chunk hand-out (one ticket per chunk vs grouped tickets) on cfg3 at N=1, plus the partition parity tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for t in 0 8; do
  B200GBM_PART_TICKETS=$t B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2n_cfg3_t$t.json 2> gpurun_out/r2n_cfg3_t$t.err
  echo "tickets=$t"; grep "split timing" gpurun_out/r2n_cfg3_t$t.err | cut -c1-300; cut -c1-160 gpurun_out/r2n_cfg3_t$t.json
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ingest_scale.py -m gpu -q -x -k "not offsets_beyond" 2>&1 | tail -3
