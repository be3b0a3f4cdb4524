#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "gpt2" 2>&1 | tail -12 > gpurun_out/t_attn2.log
tail -6 gpurun_out/t_attn2.log
timeout -k 10 200 python bench.py --steps 20 --warmup 5 --attn sdpa > gpurun_out/bench_attn_sdpa.log 2>&1
timeout -k 10 200 python bench.py --steps 20 --warmup 5 --attn native > gpurun_out/bench_attn_native.log 2>&1
for f in bench_attn_sdpa bench_attn_native; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-330 || tail -5 gpurun_out/$f.log; done
