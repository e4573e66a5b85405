#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient"): tools/gpurun_retry.sh <log> <timeout> <command>
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 120; else break; fi
done
