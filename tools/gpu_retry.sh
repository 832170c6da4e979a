#!/bin/bash
# gpu_retry.sh <timeout> <command...>: keep asking for a GPU slot until the call actually runs (the pod answers "busy"
# without charging anything when all its slots are taken)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up: no GPU slot"; exit 3
