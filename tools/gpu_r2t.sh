#!/bin/bash
# round 2, GPU call T: launch list of one XE training step on the final build
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --workload xe --steps 1 --warmup 3 --no-extras > $O/t_xe_plain.json 2> $O/t_err.log && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file $O/t_xe_launches.csv python bench.py --workload xe --steps 1 --warmup 3 --no-extras > $O/t_ncu.log 2>&1
du -sh $O
