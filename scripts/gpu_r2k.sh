#!/bin/bash
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout -k 10 500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_batchnorm.py tests/test_gpu_kernels.py -m gpu -x -q -k "conv or resnet or batchnorm or maxpool or split or deterministic" > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/tests.log
RESNET_ARM=graph timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
RESNET_ITERS=1 timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv -s 3000 -c 1000 \
    python scripts/bench_resnet_step.py > $OUT/prof.log 2>&1
python scripts/launch_shares.py $OUT/launches.csv $OUT/shares.txt | head -32
