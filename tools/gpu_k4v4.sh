#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "=== bench 100M x 512 host ingest"
timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/b100.err | grep -E '^\{' > gpurun_out/b100.json
python - <<PY
import json
d=json.load(open("gpurun_out/b100.json")); print({k:d[k] for k in ("value","ms_per_step")}, {k:d["roofline"][k] for k in ("cells_per_s","frac","k4_share_of_step")}, d["e2e"]["value"], d["e2e"]["ingest_ms"], d["clocks"])
PY
tail -3 gpurun_out/b100.err
