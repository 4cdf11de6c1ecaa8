#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <gpus> '<command>'   -- retries while the pod answers "busy" (nothing is charged)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after 40 busy answers"
