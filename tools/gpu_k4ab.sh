#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in e1 e2 e3; do
  timeout 300 build/ubench_hist_$v 20000000 1 1 > gpurun_out/ab_$v.json 2>/dev/null
  echo "== $v"; grep '"natom"' gpurun_out/ab_$v.json | head -3
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k4_hist_build_ws -s 2 -c 1 -o gpurun_out/k4v4_prof build/ubench_hist 10000000 1 1 > gpurun_out/ncu_k4v4.log 2>&1
tail -3 gpurun_out/ncu_k4v4.log
