#!/bin/bash
# block-scaled fp8 GEMM numerics + the control-plane tests that changed (mailbox, K9 in lagom); every step time-boxed
OUT=gpurun_out/mx_r2a
mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_fp8.py -x -q -k "mxfp8" 2>&1 | tail -25 | tee $OUT/mx_tests.log
timeout 120 python -m pytest tests/test_gpu_ctrl.py -x -q 2>&1 | tail -15 | tee $OUT/ctrl_tests.log
timeout 200 python -m pytest tests/test_gpu_lagom.py -x -q 2>&1 | tail -15 | tee $OUT/lagom_tests.log
