#!/bin/bash
# round 2, GPU call W: compact features (valid regions only) -- parity tests, adaptive e2e A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q --timeout 600 -rA > $O/w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/w_pytest.log
if grep -q "pytest rc=0" $O/w_pytest.log; then
  timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras --compact > $O/w_adaptive_compact.json 2> $O/w_bench.err
  timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/w_adaptive_padded.json 2>> $O/w_bench.err
  timeout 600 python bench.py --adaptive --regions 100 --batch 1024 --no-extras --compact > $O/w_adaptive_compact_b1024.json 2>> $O/w_bench.err
  timeout 600 python bench.py --adaptive --regions 100 --batch 1024 --no-extras > $O/w_adaptive_padded_b1024.json 2>> $O/w_bench.err
fi
du -sh $O
