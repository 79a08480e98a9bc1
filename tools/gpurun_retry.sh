#!/bin/bash
# Usage: tools/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] -- '<command>'
# Retries gpurun while the pod answers "busy" (exit code 3, nothing charged). The repository snapshot is taken when the
# call is admitted, so keep the tree consistent until the log says "sending".
log=$1; shift
timeout=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$timeout" "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
