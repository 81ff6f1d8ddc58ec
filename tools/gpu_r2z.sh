#!/bin/bash
# round 2, final GPU call Z: the round's final build -- whole GPU suite, smoke, bench lines of every configuration, per-shape records,
# ncu launch list, ncu --set full of the vocabulary passes and of the (opt-in) LayerNorm-epilogue GEMM
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z_smoke.log 2>&1; echo "smoke rc=$?" >> $O/z_smoke.log
timeout 900 python bench.py > $O/z_bench.json 2> $O/z_bench.err; echo "bench rc=$?" >> $O/z_bench.err
timeout 900 python bench.py --impl reference > $O/z_reference.json 2>> $O/z_bench.err
BOFI_PROFILE_DUMP=$O/z_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/z_bench_dump.json 2>> $O/z_bench.err
timeout 600 python bench.py --no-extras --no-logprobs > $O/z_nolp.json 2>> $O/z_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/z_adaptive.json 2>> $O/z_bench.err
timeout 600 python bench.py --regions 100 --batch 512 --no-extras > $O/z_r100.json 2>> $O/z_bench.err
timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/z_saic.json 2>> $O/z_bench.err
timeout 900 python bench.py --workload xe > $O/z_xe.json 2>> $O/z_bench.err
for b in 1 32 512; do
  timeout 300 python bench.py --no-extras --depth 1 --batch $b --calib s_cap --no-logprobs > $O/z_lat_$b.json 2>> $O/z_bench.err
done
BOFI_GRAPH=0 timeout 300 python tools/one_decode.py > $O/z_one_decode.log 2>&1 && \
BOFI_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/z_launches.csv python tools/one_decode.py > $O/z_ncu1.log 2>&1
BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc2_kernel -s 68 -c 2 -o /tmp/full_vocab python tools/one_decode.py > $O/z_ncu_vocab.log 2>&1
ncu -i /tmp/full_vocab.ncu-rep --page raw --csv > $O/z_full_vocabgemm.csv 2>/dev/null
BOFI_LNEPI=1 BOFI_GRAPH=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc2_ln_kernel -s 2 -c 2 -o /tmp/full_lnepi python tools/one_decode.py > $O/z_ncu_lnepi.log 2>&1
ncu -i /tmp/full_lnepi.ncu-rep --page raw --csv > $O/z_full_lnepi.csv 2>/dev/null
du -sh $O
