#!/bin/bash
# usage: gpurun_retry.sh <max_tries> <gpurun args...> : retries while the pod answers "busy" (exit 3, nothing charged)
tries=$1; shift
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i answered busy; sleeping 120 s"
  sleep 120
done
exit 3
