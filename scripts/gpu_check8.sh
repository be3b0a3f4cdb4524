#!/bin/bash
# 8-GPU box: correctness of the fused collectives at world=8, scaling, K1/K3 rooflines, ASHA trials/hour
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { timeout -k 10 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 200 29601 tests/multi/rs_adam_check.py > gpurun_out/rs_adam_check_n$N.log 2>&1
run 200 29602 tests/multi/ag_gemm_check.py > gpurun_out/ag_gemm_check_n$N.log 2>&1
run 200 29603 bench/k1_roofline.py > gpurun_out/k1_roofline_n$N.log 2>&1
NCCL_DEBUG=INFO run 240 29604 bench.py --gpus $N --steps 20 --warmup 5 --multicast > gpurun_out/bench_n$N.log 2>&1
run 240 29605 bench.py --gpus $N --steps 20 --warmup 5 --no-multicast > gpurun_out/bench_n${N}_nomc.log 2>&1
run 240 29606 bench.py --gpus $N --steps 10 --warmup 3 --impl torch > gpurun_out/bench_n${N}_torch.log 2>&1
timeout -k 10 420 python bench/hpo_asha_resnet.py --workers $N --trials 32 --steps-per-budget 10 --batch 64 > gpurun_out/hpo_asha_n$N.log 2>&1
grep -E "ok|OK|Error|assert" gpurun_out/rs_adam_check_n$N.log | tail -5
grep -E "OK|Error|assert|ag_gemm" gpurun_out/ag_gemm_check_n$N.log | tail -3 | cut -c1-1200
tail -1 gpurun_out/k1_roofline_n$N.log | cut -c1-700
grep -c "NVLS" gpurun_out/bench_n$N.log
for f in bench_n$N bench_n${N}_nomc bench_n${N}_torch; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-330; done
tail -1 gpurun_out/hpo_asha_n$N.log | cut -c1-900
