#!/bin/bash
# 8-GPU lease: BASELINE config 5 as specified (Llama-3 1B, ZeRO-3 + MXFP8 + K3-MX, LOCO ablation through lagom),
# its bf16 / fp8-forward DP comparisons, the headline bench at N=8 with guards, ASHA trials/hour (config 2)
OUT=gpurun_out/n8_r2a
mkdir -p $OUT
export MAGGY_LOG_DIR=/tmp/maggy_n8
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > $OUT/clocks.csv &
SMI=$!
timeout 420 python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 8 --mode fsdp-mx --steps 8 --batch 4 --seq 2048 --layers 2 > $OUT/cfg5_fsdp_mx.json 2> $OUT/cfg5_fsdp_mx.err
tail -c 900 $OUT/cfg5_fsdp_mx.json; tail -4 $OUT/cfg5_fsdp_mx.err
timeout 300 python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 8 --mode dp --steps 8 --batch 4 --seq 2048 --layers 2 > $OUT/cfg5_dp_bf16.json 2> $OUT/cfg5_dp_bf16.err
tail -c 700 $OUT/cfg5_dp_bf16.json; tail -2 $OUT/cfg5_dp_bf16.err
timeout 300 python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 8 --mode dp --fp8 mx --steps 8 --batch 4 --seq 2048 --layers 2 > $OUT/cfg5_dp_mxfp8.json 2> $OUT/cfg5_dp_mxfp8.err
tail -c 700 $OUT/cfg5_dp_mxfp8.json; tail -2 $OUT/cfg5_dp_mxfp8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
tail -c 1500 $OUT/bench_n8.json; tail -2 $OUT/bench_n8.err
timeout 400 python bench.py --config asha --gpus 8 --trials 48 --steps-per-budget 100 > $OUT/asha_n8.json 2> $OUT/asha_n8.err
tail -c 1200 $OUT/asha_n8.json; tail -3 $OUT/asha_n8.err
kill $SMI
