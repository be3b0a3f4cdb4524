#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_attn.log
tail -25 gpurun_out/t_attn.log
timeout -k 10 200 python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1
tail -5 gpurun_out/attn_bench.log | cut -c1-250
