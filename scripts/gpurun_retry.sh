#!/bin/bash
# gpurun with retries while the pod answers "no box or slot free right now" (exit code 3: nothing charged)
# usage: gpurun_retry.sh [gpurun args...] -- 'command'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 45 s" >&2
  sleep 45
done
exit 3
