#!/bin/bash
OUT=gpurun_out/r2i
mkdir -p $OUT
timeout -k 10 500 python -m pytest tests/test_gpu_batchnorm.py tests/test_gpu_conv.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -8 $OUT/tests.log
RESNET_ARM=graph MAGGY_B200_BN=native timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
RESNET_ARM=eager timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
BN_STEP=0 timeout 300 python scripts/bench_bn.py 2>&1 | tee $OUT/bench_bn.log | tail -9
RESNET_ITERS=1 timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv -s 3000 -c 1100 \
    python scripts/bench_resnet_step.py > $OUT/prof.log 2>&1
python scripts/launch_shares.py $OUT/launches.csv $OUT/shares.txt | head -40
