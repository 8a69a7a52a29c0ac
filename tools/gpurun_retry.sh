#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>  -- retries while the pod answers "transient" (nothing charged)
t=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 120; continue; fi
  break
done
