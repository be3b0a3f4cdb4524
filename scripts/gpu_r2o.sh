#!/bin/bash
OUT=gpurun_out/r2o
mkdir -p $OUT
PROF_WHAT=gemm timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05" -s 2 -c 6 -o $OUT/gemm_mn python scripts/prof_r2_kernels.py > $OUT/gemm.log 2>&1; echo "gemm rc=$?"
PROF_WHAT=conv timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"conv2d_" -s 4 -c 8 -o $OUT/conv python scripts/prof_r2_kernels.py > $OUT/conv.log 2>&1; echo "conv rc=$?"
PROF_WHAT=norm timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"bn_|norm_bwd_kernel_v4" -s 6 -c 8 -o $OUT/norm python scripts/prof_r2_kernels.py > $OUT/norm.log 2>&1; echo "norm rc=$?"
ls -la $OUT
