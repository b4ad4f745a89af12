#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
