#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <command...>   — retries gpurun while the pod answers "transient" / busy (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); RC=$?
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|no box\|busy"; then
    if ! echo "$OUT" | grep -q "status=ok\|status=done\|rc=0"; then sleep 90; continue; fi
  fi
  echo "$OUT" | tail -120; exit $RC
done
echo "gave up after 40 tries"; exit 3
