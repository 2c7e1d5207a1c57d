#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>' : retries while the pod answers "transient"/busy
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then
    sleep 120
    continue
  fi
  echo "$OUT" | tail -60
  echo "[retry wrapper] attempts=$i rc=$RC"
  exit $RC
done
echo "[retry wrapper] gave up"
exit 3
