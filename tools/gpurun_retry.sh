#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient": nothing charged).
# usage: tools/gpurun_retry.sh <timeout-s> <log> [--gpus N] -- <command>
TO=$1; LOG=$2; shift 2
for try in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
