#!/bin/bash
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout -k 10 600 python -m pytest tests/test_gpu_lagom.py tests/test_gpu_ctrl.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -30 $OUT/tests.log
