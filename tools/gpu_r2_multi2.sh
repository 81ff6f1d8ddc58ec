#!/bin/bash
# round 2, 2-GPU sanity call of the final build (gpurun --gpus 2): decode with the grouped pipeline + caption gather, reference arm under torchrun, XE step
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > $O/m2_decode_2.json 2>> $O/m2_err.log; echo "decode rc=$?" >> $O/m2_err.log
$TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 3 --group 1 --no-extras > $O/m2_decode_2_g1.json 2>> $O/m2_err.log
$TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --workload xe --no-extras > $O/m2_xe_2.json 2>> $O/m2_err.log
$TR --nproc-per-node 2 --master-port 29525 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/m2_reference_2.json 2>> $O/m2_err.log; echo "reference rc=$?" >> $O/m2_err.log
timeout 300 python -m pytest tests/test_sharding_gloo.py -q > $O/m2_gloo.log 2>&1
tail -5 $O/m2_err.log
du -sh $O
