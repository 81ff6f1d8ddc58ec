#!/bin/bash
# round 2, 8-GPU call of the final build (gpurun --gpus 8): headline decode (groups of two) and BASELINE config 3 (adaptive 10..100 regions,
# 4096 images over 8 GPUs) with compact features on the end-to-end leg
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --no-extras --adaptive --regions 100 --batch 512 --compact > $O/m8_adaptive_compact_8.json 2>> $O/m8_err.log; echo "rc=$?" >> $O/m8_err.log
$TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --no-extras > $O/m8_decode_8.json 2>> $O/m8_err.log; echo "rc=$?" >> $O/m8_err.log
grep -v "OMP_NUM\|\*\*\*" $O/m8_err.log | tail -5
du -sh $O
