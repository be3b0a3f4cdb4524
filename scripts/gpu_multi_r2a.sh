#!/bin/bash
# 2-GPU lease: multi-GPU kernel tests (K1, K3, K3-MX, rs_sum, sharded linears), K9 in a real lagom run, N=2 bench with guards
OUT=gpurun_out/multi_r2a
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40 | tee $OUT/multi_tests.log
timeout 200 python -m pytest tests/test_gpu_lagom.py -x -q 2>&1 | tail -15 | tee $OUT/lagom_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -c 1800 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
