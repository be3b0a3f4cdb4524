#!/bin/bash
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout -k 10 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py tests/test_gpu_batchnorm.py -m gpu -q -x > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/tests.log
RESNET_ARM=graph timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee $OUT/step.log
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-200 | tee $OUT/bench.log
