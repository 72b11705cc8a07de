#!/bin/bash
# usage: tools/grun.sh <out-file> <timeout-s> [--gpus N] -- <command>   : gpurun with retries while the pod is busy
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $out 2>&1
  if grep -q "status=transient\|rc=3" $out || grep -q "no box or slot" $out; then sleep 60; continue; fi
  break
done
