#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy/transient)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" $LOG && ! grep -q "status=ok" $LOG; then sleep 120; continue; fi
  break
done
