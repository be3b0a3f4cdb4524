#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_multi2.log
tail -25 gpurun_out/t_multi2.log
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_pair.log 2>&1
grep '^{' gpurun_out/bench_n2_pair.log | tail -1 | cut -c1-330
