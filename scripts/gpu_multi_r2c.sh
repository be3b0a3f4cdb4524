#!/bin/bash
# 2-GPU validation at the end of round 2: the multi-GPU test file, the lagom tests on 2 GPUs (K9 on the spare GPU), N=2 bench
OUT=gpurun_out/multi_r2c
mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_lagom.py -m gpu -q -x > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.log 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; grep '^{' $OUT/bench_n2.log | tail -1 | cut -c1-420
