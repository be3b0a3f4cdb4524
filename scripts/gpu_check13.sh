#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_conv.log
tail -5 gpurun_out/t_conv.log
timeout -k 10 400 python scripts/bench_conv.py --batch 64 --out gpurun_out/conv_bench_b64.json > gpurun_out/conv_bench_b64.log 2>&1
tail -16 gpurun_out/conv_bench_b64.log | cut -c1-220
