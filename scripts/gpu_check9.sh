#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or autograd or gpt2 or llama" 2>&1 | tail -15 > gpurun_out/t_mc.log
MAGGY_B200_GEMM_MC=1 timeout -k 10 200 python scripts/bench_gemm.py --tokens 32768 > gpurun_out/gemm_bench_mc1.log 2>&1
MAGGY_B200_GEMM_MC=0 timeout -k 10 200 python scripts/bench_gemm.py --tokens 32768 > gpurun_out/gemm_bench_mc0.log 2>&1
MAGGY_B200_GEMM_MC=1 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mc1.log 2>&1
MAGGY_B200_GEMM_MC=0 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mc0.log 2>&1
tail -8 gpurun_out/t_mc.log; echo MC1; tail -11 gpurun_out/gemm_bench_mc1.log | cut -c1-200; echo MC0; tail -11 gpurun_out/gemm_bench_mc0.log | cut -c1-200
for f in bench_mc1 bench_mc0; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-260; done
