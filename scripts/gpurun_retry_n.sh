#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <gpus> <timeout-seconds> '<command>'
G=$1; T=$2; shift; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1)
  echo "$OUT" | tail -60
  if ! echo "$OUT" | grep -qE "status=transient|status=busy|rc=3"; then exit 0; fi
  echo "[retry $i] no slot, sleeping 180 s"; sleep 180
done
