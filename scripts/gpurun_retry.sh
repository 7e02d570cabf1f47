#!/usr/bin/env bash
# gpurun with retries while the pod answers "no box / slot free right now" (exit code 3, nothing charged).
# usage: scripts/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ "$rc" != "3" ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt answered busy; sleeping 90 s"
  sleep 90
done
exit 3
