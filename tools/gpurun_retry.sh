#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers busy / transient (nothing charged)
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3\|rc=3"; then sleep 150; continue; fi
  break
done
