#!/bin/bash
mkdir -p gpurun_out
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or autograd or gpt2 or llama" 2>&1 | tail -15 > gpurun_out/t_pair.log
tail -3 gpurun_out/t_pair.log
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 200 python scripts/bench_gemm.py --tokens 32768 > gpurun_out/gemm_bench_pair1.log 2>&1
echo PAIR1; tail -11 gpurun_out/gemm_bench_pair1.log | cut -c1-200
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pair1.log 2>&1
for f in bench_pair1; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-300; done
