#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== full bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full2.json 2> gpurun_out/bench_full2.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_full2.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "roofline", d["roofline"], "\ncpu", d["cpu_baseline"], "\ne2e", d["e2e"], d["clocks"])
PY
tail -3 gpurun_out/bench_full2.err
