#!/bin/bash
# round 2, GPU call H: four k-blocks per ring stage in the narrow-tile GEMM (bounding loop), A/B; glancing + LN-epilogue tests
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -rA > $O/h_unit.log 2>&1; echo "unit rc=$?" >> $O/h_unit.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 -rA -k "glancing" > $O/h_glat.log 2>&1; echo "glat rc=$?" >> $O/h_glat.log
if grep -q "unit rc=0" $O/h_unit.log; then
  timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bf16_parity.py -m gpu -q --timeout 600 -rA > $O/h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/h_pytest.log
  timeout 600 python bench.py --no-extras > $O/h_kps4.json 2> $O/h_bench.err
  BOFI_KPS_SMALL=0 timeout 600 python bench.py --no-extras > $O/h_base.json 2>> $O/h_bench.err
  timeout 600 python bench.py --no-extras --depth 1 > $O/h_kps4_d1.json 2>> $O/h_bench.err
  BOFI_KPS_SMALL=0 timeout 600 python bench.py --no-extras --depth 1 > $O/h_base_d1.json 2>> $O/h_bench.err
  for b in 1 32; do
    timeout 300 python bench.py --no-extras --depth 1 --batch $b --calib s_cap --no-logprobs > $O/h_lat_$b.json 2>> $O/h_bench.err
    BOFI_KPS_SMALL=0 timeout 300 python bench.py --no-extras --depth 1 --batch $b --calib s_cap --no-logprobs > $O/h_lat_base_$b.json 2>> $O/h_bench.err
  done
  BOFI_PROFILE_DUMP=$O/h_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/h_bench_dump.json 2>> $O/h_bench.err
fi
du -sh $O
