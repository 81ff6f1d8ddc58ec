#!/bin/bash
# round 2, 2-GPU call: per-encoder-layer gradient buckets -- equality with one all-reduce, XE step times with / without
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29701 tools/check_grad_reduce.py > $O/m2b_check.log 2>&1; echo "check rc=$?" >> $O/m2b_check.log
$TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --workload xe --no-extras > $O/m2b_xe_2.json 2>> $O/m2b_err.log
tail -3 $O/m2b_check.log
