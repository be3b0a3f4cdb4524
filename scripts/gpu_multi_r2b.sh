#!/bin/bash
OUT=gpurun_out/multi_r2b
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "generic or rs_adam" 2>&1 | tail -30 | tee $OUT/generic_dp.log
timeout 200 python -m pytest tests/test_gpu_lagom.py -x -q 2>&1 | grep -E "^E |passed|failed" | head -20 | tee $OUT/lagom.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cross_entropy or parity or adam" 2>&1 | tail -5 | tee $OUT/kernels.log
