#!/bin/bash
# config 5 plumbing check on 2 GPUs (small model): LOCO ablation through lagom, ZeRO-3 + MXFP8 + K3-MX vs bf16 DP
OUT=gpurun_out/cfg5_n2
mkdir -p $OUT
export MAGGY_LOG_DIR=/tmp/maggy_cfg5
timeout 400 python bench/ablation_llama.py --model llama-small --gpus-per-trial 2 --mode fsdp-mx --steps 6 --batch 4 --seq 512 --layers 1 > $OUT/fsdp_mx.json 2> $OUT/fsdp_mx.err
tail -c 1200 $OUT/fsdp_mx.json; tail -25 $OUT/fsdp_mx.err
timeout 400 python bench/ablation_llama.py --model llama-small --gpus-per-trial 2 --mode dp --steps 6 --batch 4 --seq 512 --layers 1 > $OUT/dp.json 2> $OUT/dp.err
tail -c 1200 $OUT/dp.json; tail -5 $OUT/dp.err
timeout 200 python -m pytest tests/test_gpu_lagom.py -x -q 2>&1 | tail -5
