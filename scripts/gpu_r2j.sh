#!/bin/bash
OUT=gpurun_out/r2j
mkdir -p $OUT
timeout -k 10 500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_batchnorm.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/tests.log
export MAGGY_LOG_DIR=/tmp/asha_logs
timeout 600 python bench/hpo_asha_resnet.py --workers 1 --trials 8 --steps-per-budget 100 > $OUT/asha.log 2>&1; tail -1 $OUT/asha.log
grep -rh "TIMING" /tmp/asha_logs | head -20 | tee $OUT/timing.log
