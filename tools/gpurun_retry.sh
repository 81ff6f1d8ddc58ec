#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout_s> <command...>   -- retries while the pod answers busy (nothing is charged then)
tag=$1; shift; tmo=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $tmo -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/${tag}_call.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  echo "gpurun rc=$rc after $i tries"; exit $rc
done
echo "gave up"; exit 3
