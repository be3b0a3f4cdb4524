#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "norm or gpt2 or autograd or llama or bert" 2>&1 | tail -8 > gpurun_out/t_ln.log
tail -5 gpurun_out/t_ln.log
timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_lnres.log 2>&1
grep '^{' gpurun_out/bench_lnres.log | tail -1 | cut -c1-300
