#!/bin/bash
# round 2, GPU call L: log-prob rows on a 16-byte pitch (TMA stores in the second vocabulary pass), A/B against the dense tensor
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_sampling.py tests/test_gpu_bf16_parity.py -m gpu -q --timeout 600 -rA > $O/l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/l_pytest.log
timeout 600 python bench.py --no-extras > $O/l_padded.json 2> $O/l_bench.err
timeout 600 python bench.py --no-extras > $O/l_padded2.json 2>> $O/l_bench.err
BOFI_PROFILE_DUMP=$O/l_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/l_bench_dump.json 2>> $O/l_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/l_smoke.log 2>&1; echo "smoke rc=$?" >> $O/l_smoke.log
du -sh $O
