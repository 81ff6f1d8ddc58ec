#!/bin/bash
# round 2, GPU call R: XE training with two bounding layers (N_len = 2), regression of the training tests
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 -rA -k "two_bounding" > $O/r_nlen2.log 2>&1; echo "nlen2 rc=$?" >> $O/r_nlen2.log
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_selfcritical.py -m gpu -q --timeout 600 > $O/r_train.log 2>&1; echo "train rc=$?" >> $O/r_train.log
du -sh $O
