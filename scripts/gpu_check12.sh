#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/t_conv.log
tail -40 gpurun_out/t_conv.log
MAGGY_B200_GEMM_PAIR=0 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pair0.log 2>&1
MAGGY_B200_GEMM_PAIR=1 timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pair1.log 2>&1
for f in bench_pair0 bench_pair1; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-300; done
