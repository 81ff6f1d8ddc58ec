#!/bin/bash
# round 2, GPU call D: the bounding loop as one cluster kernel (bound_loop.cuh): parity first, then A/B benches and latencies
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_smoke.log 2>&1; echo "smoke rc=$?" >> $O/d_smoke.log
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bf16_parity.py tests/test_gpu_shapes.py -m gpu -q --timeout 300 -x -rA > $O/d_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest1.log
if grep -q "rc=0" $O/d_pytest1.log; then
  timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
fi
timeout 600 python bench.py --no-extras > $O/d_cluster.json 2> $O/d_bench.err
BOFI_BOUND_CLUSTER=0 timeout 600 python bench.py --no-extras > $O/d_graph.json 2>> $O/d_bench.err
timeout 600 python bench.py --no-extras --depth 1 > $O/d_cluster_d1.json 2>> $O/d_bench.err
timeout 600 python bench.py --no-extras --depth 2 > $O/d_cluster_d2.json 2>> $O/d_bench.err
for b in 1 32 512; do
  timeout 300 python bench.py --batch $b --depth 1 --calib s_cap --no-extras --steps 50 > $O/d_lat_$b.json 2>> $O/d_bench.err
  BOFI_BOUND_CLUSTER=0 timeout 300 python bench.py --batch $b --depth 1 --calib s_cap --no-extras --steps 50 > $O/d_lat_graph_$b.json 2>> $O/d_bench.err
done
timeout 600 python bench.py --adaptive --regions 100 --batch 512 --no-extras > $O/d_adaptive.json 2>> $O/d_bench.err
BOFI_PROFILE_DUMP=$O/d_records.csv timeout 600 python bench.py --steps 5 --no-extras > $O/d_bench_dump.json 2>> $O/d_bench.err
timeout 300 python tools/one_decode.py > $O/d_one_decode.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:bound_loop_kernel -c 1 -o /tmp/full_bl python tools/one_decode.py > $O/d_ncu_bl.log 2>&1
ncu -i /tmp/full_bl.ncu-rep --page raw --csv > $O/d_full_bound_loop.csv 2>/dev/null
du -sh $O
