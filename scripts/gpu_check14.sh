#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "graph or native_vs" 2>&1 | tail -30 > gpurun_out/t_conv2.log
tail -12 gpurun_out/t_conv2.log
for eng in graph eager native-eager; do
  timeout -k 10 300 python bench/hpo_asha_resnet.py --workers 1 --trials 12 --engine $eng > gpurun_out/asha_1gpu_$eng.log 2>&1
  echo "== $eng"; grep '^{' gpurun_out/asha_1gpu_$eng.log | tail -1 | cut -c1-420 || tail -5 gpurun_out/asha_1gpu_$eng.log
done
