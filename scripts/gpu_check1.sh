#!/bin/bash
# first contact with the GPU: kernel numerics, GEMM micro-bench, short training bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm_kmajor" 2>&1 | tail -25 > gpurun_out/t_gemm.log
echo "gemm_kmajor rc=$?" >> gpurun_out/t_gemm.log
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -60 > gpurun_out/t_all.log
timeout -k 10 300 python scripts/bench_gemm.py --tokens 32768 --out gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --impl torch > gpurun_out/bench1_torch.log 2>&1
tail -5 gpurun_out/t_gemm.log; tail -30 gpurun_out/t_all.log; cat gpurun_out/gemm_bench.log | tail -15; tail -3 gpurun_out/bench1.log; tail -3 gpurun_out/bench1_torch.log
