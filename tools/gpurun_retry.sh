#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> [--gpus N]   — retries while the pod answers "transient/busy"
T=$1; S=$2; shift 2
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" -- "bash $S" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|status=busy\|rc=None"; then sleep 150; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up after 30 tries"; echo "$OUT" | tail -5
