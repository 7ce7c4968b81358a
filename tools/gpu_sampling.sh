#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_shapes or no_usable or extreme_weights" 2>&1 | tail -40
