#!/usr/bin/env bash
# usage: gpurun_retry.sh [gpurun args...] -- 'command'   (retries while the pod answers "transient"/busy)
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
