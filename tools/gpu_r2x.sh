#!/bin/bash
# round 2, GPU call X: grouped compact submissions -- tests, adaptive e2e A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_decode.py -m gpu -q --timeout 600 -k "compact or grouped or sharded or host_slot" > $O/x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/x_pytest.log
if grep -q "pytest rc=0" $O/x_pytest.log; then
  timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras --compact > $O/x_adaptive_compact_g2.json 2> $O/x_bench.err
  timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras --compact --group 1 > $O/x_adaptive_compact_g1.json 2>> $O/x_bench.err
fi
tail -n 3 $O/x_pytest.log
