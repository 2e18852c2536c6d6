#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh LOGFILE TIMEOUT CMD...   — retries while the pod answers busy; nothing is charged for those
LOG=$1; shift; TO=$1; shift
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 45
done
exit 3
