#!/bin/bash
# gpurun with retries on "transient" answers (busy pod / back-off): scripts/gpurun_retry.sh <gpurun args...>
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt answered transient; sleeping 240 s"
    sleep 240
  else
    break
  fi
done
