#!/bin/bash
cd $GRAFT_REPO_ROOT
for b in build/ub_*; do echo "== $b"; timeout 120 $b 10000000 | grep -A8 k4_check | grep -E "mismatch|natom" ; done
