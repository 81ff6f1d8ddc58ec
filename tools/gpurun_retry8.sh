#!/bin/bash
# usage: tools/gpurun_retry8.sh <tag> <timeout_s> <command>   -- 8-GPU call, retried while the pod answers busy (nothing is charged then)
tag=$1; tmo=$2; cmd=$3
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus 8 --timeout $tmo -- "$cmd" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/${tag}_call.log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  echo "gpurun rc=$rc after $i tries"; exit $rc
done
echo "gave up"; exit 3
