#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus 8): 8-GPU decode (R = 36 and the 100-region configurations of BASELINE.json), XE training
# with the bucketed gradient all-reduce, and the box's aggregate H2D ceiling.  (1 -> 8 scaling of the headline is the driver's run.)
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/m_topo.txt 2>&1
$TR --nproc-per-node 1 --master-port 29501 tools/h2d_ceiling.py > $O/m_h2d_1.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29502 tools/h2d_ceiling.py > $O/m_h2d_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29503 tools/h2d_ceiling.py --dtype fp32 > $O/m_h2d_fp32_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29504 tools/h2d_ceiling.py --numa > $O/m_h2d_numa_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-extras > $O/m_decode_8.json 2>> $O/m_err.log
$TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --no-extras > $O/m_decode_2.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --no-extras --adaptive --regions 100 --batch 512 > $O/m_adaptive100_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --no-extras --regions 100 --batch 512 > $O/m_r100_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --workload xe > $O/m_xe_8.json 2>> $O/m_err.log
$TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --workload xe --buckets 1 > $O/m_xe_8_onebucket.json 2>> $O/m_err.log
$TR --nproc-per-node 4 --master-port 29563 bench.py --gpus 4 --workload xe > $O/m_xe_4.json 2>> $O/m_err.log
tail -5 $O/m_err.log
du -sh $O
