#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "=== cpu info"; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os;print('affinity',len(os.sched_getaffinity(0)))"; lscpu | grep -E "Model name|Socket|Thread|NUMA node\(s\)"
cat > /tmp/t_orc2.py <<'PY'
import numpy as np, time, sys, os
sys.path.insert(0, os.environ['GRAFT_REPO_ROOT'])
from oracle import oracle as O
rng = np.random.default_rng(1)
n, F = 500000, 128
X = rng.random((n, F)); y = (rng.random(n) < 0.5).astype(np.float32)
t=time.time(); ds = O.OracleDataset(X, "max_bin=255").set_field("label", y); d1=time.time()-t
b = O.OracleBooster(ds, "objective=binary num_leaves=31 verbosity=-1")
t=time.time(); b.train(5); dt=time.time()-t
hs, hc = b.hist_stats()
print("threads", O.lib().orc_num_threads(), "dataset %.2f"%d1, "5 iters %.3f"%dt, "hist cells/s %.3g"%(hc/hs))
PY
for t in 8 16 32 64 128; do OMP_NUM_THREADS=$t python /tmp/t_orc2.py; done
OMP_NUM_THREADS=32 OMP_PROC_BIND=spread OMP_PLACES=cores python /tmp/t_orc2.py
echo "=== full bench"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "=== ncu launch list (10M x 512)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --rows 10000000 --steps 2 --warmup 1 --no-cpu-baseline --ingest device > gpurun_out/ncu_launch_bench.log 2>&1; tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
echo "=== ncu full K4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k4_hist_build -c 3 -o gpurun_out/k4_prof python bench.py --rows 10000000 --steps 1 --warmup 1 --no-cpu-baseline --ingest device > gpurun_out/ncu_k4_bench.log 2>&1; tail -2 gpurun_out/ncu_k4_bench.log | cut -c1-300
ls -la gpurun_out
