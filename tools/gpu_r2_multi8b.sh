#!/bin/bash
# round 2, 8-GPU call: XE training step with per-encoder-layer gradient buckets vs one encoder bucket
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --workload xe --no-extras > $O/m8b_xe_8_layers.json 2>> $O/m8b_err.log; echo "rc=$?" >> $O/m8b_err.log
$TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --workload xe --no-extras --no-layer-buckets > $O/m8b_xe_8_onefront.json 2>> $O/m8b_err.log; echo "rc=$?" >> $O/m8b_err.log
$TR --nproc-per-node 8 --master-port 29803 tools/check_grad_reduce.py > $O/m8b_check.log 2>&1; echo "check rc=$?" >> $O/m8b_check.log
grep -v "OMP_NUM\|\*\*\*" $O/m8b_err.log | tail -4; tail -2 $O/m8b_check.log
