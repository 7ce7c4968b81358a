#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sklearn" 2>&1 | tail -3
