#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> '<command>'   -- retries while the pod answers busy / transient (exit 3 or "transient")
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc"
