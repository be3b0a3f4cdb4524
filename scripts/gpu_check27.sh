#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cross_entropy" 2>&1 | tail -5
bash scripts/gpu_check26.sh
