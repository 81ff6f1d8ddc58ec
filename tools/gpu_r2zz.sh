#!/bin/bash
# round 2, closing check of the shipped library (rebuilt from the committed sources): kernel unit tests, decode parity, smoke
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q --timeout 600 > $O/zz_pytest.log 2>&1; echo "pytest rc=$?" >> $O/zz_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/zz_smoke.log 2>&1; echo "smoke rc=$?" >> $O/zz_smoke.log
tail -n 2 $O/zz_pytest.log $O/zz_smoke.log
