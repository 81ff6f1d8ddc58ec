#!/bin/bash
# round 2, GPU call S: XE step with two bounding layers, grouped-pipeline tests again after the clean-up
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --workload xe --n-len 2 --no-extras > $O/s_xe_nlen2.json 2> $O/s_bench.err
timeout 900 python bench.py --workload xe --no-extras > $O/s_xe_nlen1.json 2>> $O/s_bench.err
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 300 -k "grouped or sharded" > $O/s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s_pytest.log
du -sh $O
