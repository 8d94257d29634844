#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$OUT" | tail -80
  if ! echo "$OUT" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i] transient, sleeping 150 s"; sleep 150
done
