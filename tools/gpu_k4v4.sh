#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ./build/ubench_hist 10000000 1 1 > gpurun_out/ubench_v4.json 2> gpurun_out/ubench_v4.err
tail -12 gpurun_out/ubench_v4.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
echo "=== bench 20M x 512"
timeout 600 python bench.py --rows 20000000 --steps 6 --warmup 3 --ingest device 2>gpurun_out/b20.err | grep -E '^\{' > gpurun_out/b20.json
python - <<PY
import json
d=json.load(open("gpurun_out/b20.json")); print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"])
PY
