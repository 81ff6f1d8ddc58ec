#!/bin/bash
# round 2, last GPU call (Z3 / Z4): the tree as committed -- whole GPU suite, smoke, bench, reference arm
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/z4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/z4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z4_smoke.log 2>&1; echo "smoke rc=$?" >> $O/z4_smoke.log
timeout 900 python bench.py > $O/z4_bench.json 2> $O/z4_bench.err; echo "bench rc=$?" >> $O/z4_bench.err
timeout 900 python bench.py --impl reference > $O/z4_reference.json 2>> $O/z4_bench.err
tail -n 2 $O/z4_pytest.log $O/z4_smoke.log
