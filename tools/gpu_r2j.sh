#!/bin/bash
# round 2, GPU call J: bounding-loop graph on a high-priority stream, A/B at 1-4 batches in flight
mkdir -p gpurun_out
O=gpurun_out
for d in 3 2 4; do
  timeout 600 python bench.py --no-extras --depth $d > $O/j_prio_d$d.json 2>> $O/j_bench.err
  BOFI_BOUND_PRIO=0 timeout 600 python bench.py --no-extras --depth $d > $O/j_base_d$d.json 2>> $O/j_bench.err
done
timeout 600 python bench.py --no-extras --no-logprobs > $O/j_prio_nolp.json 2>> $O/j_bench.err
BOFI_BOUND_PRIO=0 timeout 600 python bench.py --no-extras --no-logprobs > $O/j_base_nolp.json 2>> $O/j_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/j_prio_adaptive.json 2>> $O/j_bench.err
BOFI_BOUND_PRIO=0 timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/j_base_adaptive.json 2>> $O/j_bench.err
timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/j_prio_saic.json 2>> $O/j_bench.err
BOFI_BOUND_PRIO=0 timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/j_base_saic.json 2>> $O/j_bench.err
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 300 -x > $O/j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j_pytest.log
du -sh $O
