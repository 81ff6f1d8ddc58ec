#!/bin/bash
# round 2, final GPU call Z2: the final build with the grouped pipeline as bench default -- whole GPU suite, smoke, bench lines
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/z2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/z2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/z2_smoke.log
timeout 900 python bench.py > $O/z2_bench.json 2> $O/z2_bench.err; echo "bench rc=$?" >> $O/z2_bench.err
timeout 900 python bench.py --impl reference > $O/z2_reference.json 2>> $O/z2_bench.err
timeout 600 python bench.py --no-extras > $O/z2_bench2.json 2>> $O/z2_bench.err
timeout 600 python bench.py --no-extras --group 1 > $O/z2_g1.json 2>> $O/z2_bench.err
timeout 600 python bench.py --no-extras --no-logprobs > $O/z2_nolp.json 2>> $O/z2_bench.err
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/z2_adaptive.json 2>> $O/z2_bench.err
timeout 600 python bench.py --regions 100 --batch 512 --no-extras > $O/z2_r100.json 2>> $O/z2_bench.err
timeout 600 python bench.py --mode SAIC --no-extras --no-logprobs > $O/z2_saic.json 2>> $O/z2_bench.err
du -sh $O
