#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export B200GBM_L2_FETCH=32
timeout 120 build/ubench_hist_tma 40000000 1 1 > gpurun_out/tma.json 2>gpurun_out/tma.err
grep -E 'k4_check|"natom"' gpurun_out/tma.json; tail -2 gpurun_out/tma.err
