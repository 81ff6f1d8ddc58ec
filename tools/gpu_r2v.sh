#!/bin/bash
# round 2, GPU call V: the whole GPU suite three times in a row (flakiness check of the final build)
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3; do
  timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > $O/v_pytest_$i.log 2>&1; echo "pytest rc=$?" >> $O/v_pytest_$i.log
done
tail -n 3 $O/v_pytest_*.log
