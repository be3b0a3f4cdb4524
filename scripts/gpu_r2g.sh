#!/bin/bash
OUT=gpurun_out/r2g
mkdir -p $OUT
timeout -k 10 400 python -m pytest tests/test_gpu_batchnorm.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/tests.log
timeout -k 10 400 python scripts/bench_bn.py $OUT/bench_bn.json 2>&1 | tee $OUT/bench_bn.log | tail -14
