#!/bin/bash
mkdir -p gpurun_out
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|nvjet|cutlass|gemm" -s 3 -c 5 \
   -o gpurun_out/prof_gemm_pair2 python scripts/prof_gemm2.py > gpurun_out/prof_gemm_pair2.log 2>&1
tail -5 gpurun_out/prof_gemm_pair2.log
