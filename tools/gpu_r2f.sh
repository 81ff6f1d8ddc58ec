#!/bin/bash
# round 2, GPU call F: vocabulary pass 2 with aligned 16-byte stores (A/B), the prefetching drop-in leg, memcheck of every path
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bf16_parity.py tests/test_gpu_sampling.py tests/test_gpu_shapes.py -m gpu -q --timeout 300 -rA > $O/f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f_pytest.log
timeout 600 python bench.py --no-extras > $O/f_vec.json 2> $O/f_bench.err
BOFI_VOCAB_VEC=0 timeout 600 python bench.py --no-extras > $O/f_scalar.json 2>> $O/f_bench.err
timeout 900 python bench.py > $O/f_bench.json 2>> $O/f_bench.err; echo "bench rc=$?" >> $O/f_bench.err
timeout 300 python tools/sanitize_small.py > $O/f_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/f_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/f_memcheck.log
tail -c 3000 $O/f_memcheck.log > $O/f_memcheck_tail.log
BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc2_kernel -s 68 -c 2 -o /tmp/full_vocab python tools/one_decode.py > $O/f_ncu_vocab.log 2>&1
ncu -i /tmp/full_vocab.ncu-rep --page raw --csv > $O/f_full_vocabgemm.csv 2>/dev/null
du -sh $O
