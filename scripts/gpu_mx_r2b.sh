#!/bin/bash
OUT=gpurun_out/mx_r2b
mkdir -p $OUT
timeout 300 python scripts/bench_gemm_fp8.py 2>&1 | tee $OUT/gemm_fp8_mx_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_mxfp8_kernel -s 2 -c 1 -o $OUT/prof_gemm_mxfp8 python scripts/prof_mx.py > $OUT/ncu.log 2>&1
ls -la $OUT
