#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_fp8.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_fp8.log
tail -25 gpurun_out/t_fp8.log
timeout -k 10 200 python scripts/bench_gemm_fp8.py > gpurun_out/gemm_fp8_bench.log 2>&1
tail -6 gpurun_out/gemm_fp8_bench.log | cut -c1-250
