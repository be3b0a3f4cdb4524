#!/bin/bash
# 8-GPU box: correctness of the fused collectives at world=8, scaling, K1/K3 rooflines, HPO configs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
R=$GRAFT_REPO_ROOT
run() { timeout -k 10 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 150 29601 tests/multi/rs_adam_check.py > gpurun_out/rs_adam_check_n$N.log 2>&1
run 150 29602 tests/multi/ag_gemm_check.py > gpurun_out/ag_gemm_check_n$N.log 2>&1
run 150 29603 bench/k1_roofline.py > gpurun_out/k1_roofline_n$N.log 2>&1
NCCL_DEBUG=INFO run 200 29604 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
run 200 29606 bench.py --gpus $N --steps 6 --warmup 3 --impl torch > gpurun_out/bench_n${N}_torch.log 2>&1
cd /tmp && rm -rf hpo_run && mkdir hpo_run && cd hpo_run
timeout -k 10 300 python $R/bench/hpo_asha_resnet.py --workers $N --trials 32 --steps-per-budget 10 --batch 64 > $R/gpurun_out/hpo_asha_n$N.log 2>&1
timeout -k 10 240 python $R/bench/hpo_gp_bert.py --gpus-per-trial 2 --trials 8 --steps 30 --batch 16 > $R/gpurun_out/hpo_gp_bert_n$N.log 2>&1
timeout -k 10 300 python $R/bench/ablation_llama.py --model llama3-1b --steps 8 --batch 2 --seq 2048 --layers 2 > $R/gpurun_out/ablation_llama_n$N.log 2>&1
cd $R
grep -E "ok|OK|Error|assert" gpurun_out/rs_adam_check_n$N.log | tail -4
grep -E "OK|Error|assert|ag_gemm" gpurun_out/ag_gemm_check_n$N.log | tail -3 | cut -c1-1200
tail -1 gpurun_out/k1_roofline_n$N.log | cut -c1-700
echo "NVLS lines: $(grep -c NVLS gpurun_out/bench_n$N.log)"
for f in bench_n$N bench_n${N}_torch; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-330; done
for f in hpo_asha_n$N hpo_gp_bert_n$N ablation_llama_n$N; do tail -1 gpurun_out/$f.log | cut -c1-700; done
