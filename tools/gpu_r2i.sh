#!/bin/bash
# round 2, GPU call I: micro-experiments on the default build -- bounding head with 16 rows per CTA, LayerNorm with two rows per warp,
# four batches in flight
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 300 -x > $O/i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest.log
timeout 600 python bench.py --no-extras > $O/i_base.json 2> $O/i_bench.err
BOFI_HEAD_ROWS=16 timeout 600 python bench.py --no-extras > $O/i_head16.json 2>> $O/i_bench.err
BOFI_LN_ROWS=2 timeout 600 python bench.py --no-extras > $O/i_ln2.json 2>> $O/i_bench.err
BOFI_HEAD_ROWS=16 BOFI_LN_ROWS=2 timeout 600 python bench.py --no-extras > $O/i_both.json 2>> $O/i_bench.err
timeout 600 python bench.py --no-extras --depth 4 > $O/i_depth4.json 2>> $O/i_bench.err
timeout 600 python bench.py --no-extras --depth 2 > $O/i_depth2.json 2>> $O/i_bench.err
BOFI_HEAD_ROWS=16 BOFI_LN_ROWS=2 timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 300 -x -k "golden or fast_len" > $O/i_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest2.log
BOFI_HEAD_ROWS=16 BOFI_LN_ROWS=2 BOFI_PROFILE_DUMP=$O/i_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/i_bench_dump.json 2>> $O/i_bench.err
du -sh $O
