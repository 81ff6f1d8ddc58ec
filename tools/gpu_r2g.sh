#!/bin/bash
# round 2, GPU call G: the residual GEMM with the LayerNorm epilogue (gemm_tc2_ln.cuh) -- unit test, decode parity, A/B bench;
# glancing-training tests
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -rA -k "residual_layernorm or inplace_residual" > $O/g_unit.log 2>&1; echo "unit rc=$?" >> $O/g_unit.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 -rA -k "glancing" > $O/g_glat.log 2>&1; echo "glat rc=$?" >> $O/g_glat.log
if grep -q "unit rc=0" $O/g_unit.log; then
  timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bf16_parity.py tests/test_gpu_shapes.py tests/test_gpu_sampling.py -m gpu -q --timeout 600 -rA > $O/g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/g_pytest.log
  timeout 600 python bench.py --no-extras > $O/g_lnepi.json 2> $O/g_bench.err
  BOFI_LNEPI=0 timeout 600 python bench.py --no-extras > $O/g_base.json 2>> $O/g_bench.err
  timeout 600 python bench.py --no-extras --no-logprobs > $O/g_lnepi_nolp.json 2>> $O/g_bench.err
  BOFI_LNEPI=0 timeout 600 python bench.py --no-extras --no-logprobs > $O/g_base_nolp.json 2>> $O/g_bench.err
  timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/g_adaptive.json 2>> $O/g_bench.err
  BOFI_LNEPI=0 timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/g_adaptive_base.json 2>> $O/g_bench.err
  BOFI_PROFILE_DUMP=$O/g_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/g_bench_dump.json 2>> $O/g_bench.err
fi
du -sh $O
