#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_ctrl.py -m gpu -q -x --timeout 60 2>&1 | tail -30 > gpurun_out/t_ctrl.log
timeout -k 5 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mn_major or gpt2_step" 2>&1 | tail -15 > gpurun_out/t_mn.log
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi/rs_adam_check.py > gpurun_out/rs_adam_check.log 2>&1
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tests/multi/ag_gemm_check.py > gpurun_out/ag_gemm_check.log 2>&1
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench/k1_roofline.py > gpurun_out/k1_roofline.log 2>&1
timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1
tail -12 gpurun_out/t_ctrl.log; tail -8 gpurun_out/t_mn.log; grep -E "maggy_b200|Error|assert|ok|OK" gpurun_out/rs_adam_check.log | tail -12; grep -E "maggy_b200|Error|assert|OK|ag_gemm" gpurun_out/ag_gemm_check.log | tail -8 | cut -c1-1500; tail -2 gpurun_out/k1_roofline.log | cut -c1-900; tail -1 gpurun_out/bench_n1.log | cut -c1-300
