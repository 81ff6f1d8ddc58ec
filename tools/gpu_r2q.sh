#!/bin/bash
# round 2, GPU call Q: host submissions staged on per-slot copy streams (underneath the slot's own decode), depth x group sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_shapes.py -m gpu -q --timeout 600 -rA -k "sharded or grouped or host_slot or varlen or non_prefix" > $O/q_pytest.log 2>&1; echo "pytest rc=$?" >> $O/q_pytest.log
if grep -q "pytest rc=0" $O/q_pytest.log; then
  for dg in "3 1" "3 2" "2 2" "2 3" "3 3" "2 4" "4 2" "1 3"; do
    set -- $dg
    timeout 600 python bench.py --no-extras --depth $1 --group $2 > $O/q_d$1_g$2.json 2>> $O/q_bench.err
  done
  timeout 600 python bench.py --no-extras --depth 3 --group 3 --adaptive --regions 100 --batch 512 > $O/q_d3_g3_adaptive.json 2>> $O/q_bench.err
  timeout 600 python bench.py --no-extras --depth 3 --group 1 --adaptive --regions 100 --batch 512 > $O/q_d3_g1_adaptive.json 2>> $O/q_bench.err
fi
du -sh $O
