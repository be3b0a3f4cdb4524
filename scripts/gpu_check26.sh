#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 2 --steps 6 --batch 4 --seq 2048 --layers 2 > gpurun_out/ablation_llama_n2.log 2>&1
echo "== ablation"; grep '^{' gpurun_out/ablation_llama_n2.log | tail -1 | cut -c1-700 || tail -8 gpurun_out/ablation_llama_n2.log
tail -3 gpurun_out/ablation_llama_n2.log | cut -c1-300
