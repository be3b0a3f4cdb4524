#!/bin/bash
mkdir -p gpurun_out
# 1. CTA-pair GEMM correctness (guarded by short timeouts: a wrong barrier protocol hangs)
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or autograd or gpt2 or llama" 2>&1 | tail -15 > gpurun_out/t_pair.log
echo "PAIR tests rc=$?"; tail -6 gpurun_out/t_pair.log
# 2. micro bench pair vs single
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 200 python scripts/bench_gemm.py --tokens 32768 > gpurun_out/gemm_bench_pair1.log 2>&1
MAGGY_B200_GEMM_PAIR=0 timeout -k 10 200 python scripts/bench_gemm.py --tokens 32768 > gpurun_out/gemm_bench_pair0.log 2>&1
echo PAIR1; tail -11 gpurun_out/gemm_bench_pair1.log | cut -c1-200; echo PAIR0; tail -11 gpurun_out/gemm_bench_pair0.log | cut -c1-200
# 3. full GPU suite with defaults (covers the new CE kernel + test_gpu_lagom)
timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t_full.log
tail -8 gpurun_out/t_full.log
# 4. end-to-end bench
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pair1.log 2>&1
MAGGY_B200_GEMM_PAIR=0 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pair0.log 2>&1
for f in bench_pair1 bench_pair0; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-300; done
