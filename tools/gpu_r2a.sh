#!/bin/bash
# round 2, GPU call A: tests, smoke, bench legs, per-shape records, ncu launch list + memory-bound kernel captures
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/a_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.log 2>&1; echo "smoke rc=$?" >> $O/a_smoke.log
timeout 900 python bench.py > $O/a_bench.json 2> $O/a_bench.err; echo "bench rc=$?" >> $O/a_bench.err
BOFI_PROFILE_DUMP=$O/a_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/a_bench_dump.json 2>> $O/a_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/a_adaptive.json 2>> $O/a_bench.err
BOFI_VARLEN=0 timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/a_adaptive_padded.json 2>> $O/a_bench.err
timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/a_saic.json 2>> $O/a_bench.err
timeout 600 python bench.py --no-extras --no-logprobs > $O/a_nolp.json 2>> $O/a_bench.err
timeout 600 python bench.py --no-extras --depth 1 > $O/a_depth1.json 2>> $O/a_bench.err
BOFI_VOCAB_FUSED=0 timeout 600 python bench.py --no-extras > $O/a_bench_unfused.json 2>> $O/a_bench.err
BOFI_VOCAB_FUSED=0 timeout 600 python bench.py --no-extras --no-logprobs > $O/a_nolp_unfused.json 2>> $O/a_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/a_reference.json 2>> $O/a_bench.err
timeout 600 python bench.py --workload xe > $O/a_xe.json 2>> $O/a_bench.err
BOFI_GRAPH=0 timeout 300 python tools/one_decode.py > $O/a_one_decode.log 2>&1 && \
BOFI_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/a_launches.csv python tools/one_decode.py > $O/a_ncu1.log 2>&1
for k in attention_mma_kernel attention_row_bf16 bound_head bound_self_attn vocab_merge layernorm_kernel gemm_tc2_kernel; do
  BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:$k -s 2 -c 2 -o /tmp/full_$k python tools/one_decode.py > $O/a_ncu_$k.log 2>&1
  ncu -i /tmp/full_$k.ncu-rep --page raw --csv > $O/a_full_$k.csv 2>/dev/null
done
# the two vocabulary passes of the fused projection (the last two gemm_tc2 launches of a decode)
BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc2_kernel -s 83 -c 3 -o /tmp/full_vocab python tools/one_decode.py > $O/a_ncu_vocabgemm.log 2>&1
ncu -i /tmp/full_vocab.ncu-rep --page raw --csv > $O/a_full_vocabgemm.csv 2>/dev/null
du -sh $O
ls -la $O | tail -30
