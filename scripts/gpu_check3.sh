#!/bin/bash
# 2-GPU box: fused rs+adam correctness, device control plane, 1- and 2-GPU training bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout -k 10 300 python -m pytest tests/test_gpu_ctrl.py -m gpu -q -x --timeout 120 2>&1 | tail -30 > gpurun_out/t_ctrl.log
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/t_multi.log
timeout -k 10 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
NCCL_DEBUG=WARN timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-multicast > gpurun_out/bench_n2_nomc.log 2>&1
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 3 --impl torch > gpurun_out/bench_n2_torch.log 2>&1
tail -15 gpurun_out/t_ctrl.log; tail -25 gpurun_out/t_multi.log; for f in bench_n1 bench_n2 bench_n2_nomc bench_n2_torch; do tail -2 gpurun_out/$f.log | cut -c1-700; done
