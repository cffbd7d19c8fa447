#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged). Usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt: no slot, sleeping 150 s"
  sleep 150
done
exit 3
