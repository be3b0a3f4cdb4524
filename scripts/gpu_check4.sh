#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 python scripts/debug_k9.py > gpurun_out/debug_k9.log 2>&1
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi/rs_adam_check.py > gpurun_out/rs_adam_check.log 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --durations=8 -k "not kmajor" 2>&1 | tail -25 > gpurun_out/t_kernels.log
timeout -k 10 300 python scripts/bench_gemm.py --tokens 32768 --out gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1
timeout -k 10 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1
cat gpurun_out/debug_k9.log | tail -20; grep -v "^\[W\|^W0\|NCCL\|^$" gpurun_out/rs_adam_check.log | tail -25; tail -22 gpurun_out/t_kernels.log; tail -12 gpurun_out/gemm_bench.log; tail -1 gpurun_out/bench_n1.log | cut -c1-400
