#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cross_entropy or colsum or gpt2" 2>&1 | tail -8 > gpurun_out/t_ce.log
tail -4 gpurun_out/t_ce.log
timeout -k 10 200 python scripts/bench_bw_kernels.py gpurun_out/bw_kernels.json > gpurun_out/bw_kernels.log 2>&1
grep "cross_entropy\|device copy" gpurun_out/bw_kernels.log
timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ce.log 2>&1
grep '^{' gpurun_out/bench_ce.log | tail -1 | cut -c1-300
