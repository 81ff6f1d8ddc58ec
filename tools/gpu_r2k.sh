#!/bin/bash
# round 2, GPU call K: dynamic tile scheduling (cluster launch control) in the pair GEMM, alone and with the high-priority bounding stream
mkdir -p gpurun_out
O=gpurun_out
BOFI_GEMM_DYN=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -rA -k "tcgen05 or inplace_residual" > $O/k_unit.log 2>&1; echo "unit rc=$?" >> $O/k_unit.log
if grep -q "unit rc=0" $O/k_unit.log; then
  BOFI_GEMM_DYN=1 timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bf16_parity.py tests/test_gpu_shapes.py -m gpu -q --timeout 600 > $O/k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest.log
  timeout 600 python bench.py --no-extras > $O/k_base.json 2> $O/k_bench.err
  BOFI_GEMM_DYN=1 timeout 600 python bench.py --no-extras > $O/k_dyn.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 BOFI_BOUND_PRIO=1 timeout 600 python bench.py --no-extras > $O/k_dyn_prio.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 timeout 600 python bench.py --no-extras --depth 1 > $O/k_dyn_d1.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 timeout 600 python bench.py --no-extras --depth 2 > $O/k_dyn_d2.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 BOFI_BOUND_PRIO=1 timeout 600 python bench.py --no-extras --depth 2 > $O/k_dyn_prio_d2.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 timeout 600 python bench.py --no-extras --depth 4 > $O/k_dyn_d4.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 timeout 600 python bench.py --no-extras --no-logprobs > $O/k_dyn_nolp.json 2>> $O/k_bench.err
  BOFI_GEMM_DYN=1 BOFI_PROFILE_DUMP=$O/k_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/k_bench_dump.json 2>> $O/k_bench.err
fi
du -sh $O
