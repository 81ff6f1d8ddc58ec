#!/bin/bash
# round 2, GPU call B: tests on the new kernels (persistent attention, lighter vocabulary epilogue, self-critical tape), A/B benches
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest.log
timeout 600 python tools/debug_sc.py > $O/b_debug_sc.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/b_smoke.log 2>&1; echo "smoke rc=$?" >> $O/b_smoke.log
timeout 900 python bench.py > $O/b_bench.json 2> $O/b_bench.err; echo "bench rc=$?" >> $O/b_bench.err
BOFI_PROFILE_DUMP=$O/b_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/b_bench_dump.json 2>> $O/b_bench.err
timeout 600 python bench.py --no-extras --no-logprobs > $O/b_nolp.json 2>> $O/b_bench.err
timeout 600 python bench.py --no-extras --depth 1 > $O/b_depth1.json 2>> $O/b_bench.err
BOFI_DEBUG_MAX_BOUND_STEPS=0 timeout 600 python bench.py --no-extras > $O/b_nobound.json 2>> $O/b_bench.err
BOFI_DEBUG_MAX_BOUND_STEPS=0 timeout 600 python bench.py --no-extras --depth 1 > $O/b_nobound_depth1.json 2>> $O/b_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/b_adaptive.json 2>> $O/b_bench.err
timeout 600 python bench.py --regions 100 --batch 512 --no-extras > $O/b_r100.json 2>> $O/b_bench.err
for b in 1 32 512; do timeout 300 python bench.py --batch $b --depth 1 --calib s_cap --no-extras --steps 50 > $O/b_lat_$b.json 2>> $O/b_bench.err; done
BOFI_GRAPH=0 timeout 300 python tools/one_decode.py > $O/b_one_decode.log 2>&1 && \
BOFI_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/b_launches.csv python tools/one_decode.py > $O/b_ncu1.log 2>&1
BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:attention_mma_kernel -s 2 -c 2 -o /tmp/full_att python tools/one_decode.py > $O/b_ncu_att.log 2>&1
ncu -i /tmp/full_att.ncu-rep --page raw --csv > $O/b_full_attention_mma.csv 2>/dev/null
BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc2_kernel -s 68 -c 2 -o /tmp/full_vocab python tools/one_decode.py > $O/b_ncu_vocab.log 2>&1
ncu -i /tmp/full_vocab.ncu-rep --page raw --csv > $O/b_full_vocabgemm.csv 2>/dev/null
du -sh $O
