#!/bin/bash
# round 2, GPU call N: batch-size sweep of the decode (how much of the step is the bounding loop's fixed cost), XE step with the eager bar
mkdir -p gpurun_out
O=gpurun_out
for b in 2048 3072 4096; do
  timeout 600 python bench.py --no-extras --batch $b --steps 10 > $O/n_b${b}_d3.json 2>> $O/n_bench.err
  timeout 600 python bench.py --no-extras --batch $b --steps 10 --depth 2 > $O/n_b${b}_d2.json 2>> $O/n_bench.err
done
timeout 600 python bench.py --no-extras --batch 512 > $O/n_b512_d3.json 2>> $O/n_bench.err
timeout 900 python bench.py --workload xe > $O/n_xe.json 2>> $O/n_bench.err
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_selfcritical.py -m gpu -q --timeout 600 > $O/n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/n_pytest.log
du -sh $O
