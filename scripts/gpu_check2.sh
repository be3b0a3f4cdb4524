#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_all.log
timeout -k 10 300 python scripts/bench_gemm.py --tokens 32768 --out gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1
tail -30 gpurun_out/t_all.log; cat gpurun_out/gemm_bench.log | tail -15; tail -3 gpurun_out/bench1.log
