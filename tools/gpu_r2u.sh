#!/bin/bash
# round 2, GPU call U: in-kernel stall accounting of the tcgen05 GEMM on the final build (stall-counter library)
mkdir -p gpurun_out
O=gpurun_out
BOFI_LIB_PATH=boficap_b200/lib/libbofi_b200_prof.so timeout 600 python tools/gemm_stalls.py > $O/u_gemm_stalls.txt 2> $O/u_err.log; echo "rc=$?" >> $O/u_err.log
BOFI_LIB_PATH=boficap_b200/lib/libbofi_b200_prof.so timeout 600 python tools/gemm_stalls.py --batch 2048 > $O/u_gemm_stalls_b2048.txt 2>> $O/u_err.log
du -sh $O
