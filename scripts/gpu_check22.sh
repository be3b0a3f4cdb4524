#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 2 --steps 6 --batch 4 --seq 2048 --layers 2 > gpurun_out/ablation_llama_n2.log 2>&1
echo "== ablation"; grep '^{' gpurun_out/ablation_llama_n2.log | tail -1 | cut -c1-600 || tail -8 gpurun_out/ablation_llama_n2.log
for eng in graph eager; do
  timeout -k 10 300 python bench/hpo_asha_resnet.py --trials 24 --engine $eng > gpurun_out/asha_n2_$eng.log 2>&1
  echo "== asha $eng"; grep '^{' gpurun_out/asha_n2_$eng.log | tail -1 | cut -c1-420 || tail -5 gpurun_out/asha_n2_$eng.log
done
timeout -k 10 300 python -m pytest tests/test_gpu_lagom.py -m gpu -q -x 2>&1 | tail -5
