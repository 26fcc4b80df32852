#!/bin/bash
# usage: tools/grun.sh <log> [gpurun options] -- <command>    -- gpurun with retries while the pod has no free slot (nothing is charged for those)
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy" "$LOG"; then sleep 90; continue; fi
  break
done
