#!/bin/bash
# round 2, GPU call P: batches of a group decoded by one library call (bofi_set_shard / bofi_stage_part), tests + depth x group sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_shapes.py -m gpu -q --timeout 600 -rA -k "sharded or grouped or host_slot or golden" > $O/p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/p_pytest.log
if grep -q "pytest rc=0" $O/p_pytest.log; then
  for dg in "3 1" "3 2" "2 2" "2 3" "3 3" "1 3" "2 4"; do
    set -- $dg
    timeout 600 python bench.py --no-extras --depth $1 --group $2 > $O/p_d$1_g$2.json 2>> $O/p_bench.err
  done
  timeout 600 python bench.py --no-extras --depth 2 --group 2 --no-logprobs > $O/p_d2_g2_nolp.json 2>> $O/p_bench.err
  timeout 600 python bench.py --no-extras --depth 2 --group 2 --adaptive --regions 100 --batch 512 > $O/p_d2_g2_adaptive.json 2>> $O/p_bench.err
fi
du -sh $O
