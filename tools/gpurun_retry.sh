#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <log-name> <command...>   -- retries while the pod answers "busy" (rc 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /root/repo/gpurun_out/$LOG 2>&1
  rc=$?
  echo "attempt $i rc=$rc $(date)" >> /root/repo/gpurun_out/retry_attempts.log
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo "rc=$rc" >> /root/repo/gpurun_out/$LOG
