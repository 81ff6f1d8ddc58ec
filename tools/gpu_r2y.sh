#!/bin/bash
# round 2, GPU call Y: LayerNorm-epilogue GEMM with an L2 prefetch of the residual panel -- per-shape event records, unit test
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "residual_layernorm" > $O/y_unit.log 2>&1; echo "unit rc=$?" >> $O/y_unit.log
BOFI_LNEPI=1 BOFI_PROFILE_DUMP=$O/y_records.csv timeout 600 python bench.py --steps 5 --no-extras --group 1 > $O/y_bench_dump.json 2> $O/y_bench.err
tail -n 2 $O/y_unit.log
