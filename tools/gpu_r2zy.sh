#!/bin/bash
# round 2: self-critical tape pass with two bounding layers (N_len = 2), training / self-critical regression
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_selfcritical.py tests/test_gpu_train.py -m gpu -q --timeout 600 > $O/zy_pytest.log 2>&1; echo "pytest rc=$?" >> $O/zy_pytest.log
tail -n 12 $O/zy_pytest.log
