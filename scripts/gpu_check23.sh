#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_n1.log 2>&1
echo "== ref n1"; grep '^{' gpurun_out/bench_ref_n1.log | tail -1 | cut -c1-900 || tail -5 gpurun_out/bench_ref_n1.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_ref_n2.log 2>&1
echo "== ref n2"; grep '^{' gpurun_out/bench_ref_n2.log | tail -1 | cut -c1-1200 || tail -8 gpurun_out/bench_ref_n2.log
bash scripts/gpu_check22.sh
