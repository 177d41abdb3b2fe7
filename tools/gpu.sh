#!/bin/bash
# usage: tools/gpu.sh <timeout_s> <tag> '<command>'  -- retries while the pod answers "busy" (exit 3), logs to gpurun_out/<tag>_call.log
T=$1; TAG=$2; shift 2
mkdir -p /root/repo/gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /root/repo/gpurun_out/${TAG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
