#!/bin/bash
# partition kernel: sweep of the ticket grouping (chunks per ticket) on cfg3 at N=1
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for t in 1 2 4 32; do
  B200GBM_PART_TICKETS=$t B200GBM_SPLIT_TIMING=1 timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-verify --ingest device > gpurun_out/r2q_cfg3_t$t.json 2> gpurun_out/r2q_cfg3_t$t.err
  echo "tickets=$t"; grep "split timing" gpurun_out/r2q_cfg3_t$t.err | cut -c60-300
done
