#!/bin/bash
# round 2, GPU call E: the round's final build -- whole GPU suite, smoke, the bench lines of every configuration, profiles
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/e_smoke.log 2>&1; echo "smoke rc=$?" >> $O/e_smoke.log
timeout 900 python bench.py > $O/e_bench.json 2> $O/e_bench.err; echo "bench rc=$?" >> $O/e_bench.err
timeout 900 python bench.py --impl reference > $O/e_reference.json 2>> $O/e_bench.err
BOFI_PROFILE_DUMP=$O/e_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/e_bench_dump.json 2>> $O/e_bench.err
timeout 600 python bench.py --no-extras --no-logprobs > $O/e_nolp.json 2>> $O/e_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/e_adaptive.json 2>> $O/e_bench.err
BOFI_VARLEN=0 timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/e_adaptive_padded.json 2>> $O/e_bench.err
timeout 600 python bench.py --regions 100 --batch 512 --no-extras > $O/e_r100.json 2>> $O/e_bench.err
timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/e_saic.json 2>> $O/e_bench.err
timeout 600 python bench.py --workload xe > $O/e_xe.json 2>> $O/e_bench.err
BOFI_GRAPH=0 timeout 300 python tools/one_decode.py > $O/e_one_decode.log 2>&1 && \
BOFI_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/e_launches.csv python tools/one_decode.py > $O/e_ncu1.log 2>&1
du -sh $O
