#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <script> [log]   -- retries while the pool answers busy (exit code 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "${3:-/tmp/gpu_retry.log}" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
