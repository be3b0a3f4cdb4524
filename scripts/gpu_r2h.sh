#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
RESNET_ARM=graph MAGGY_B200_BN=native timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
RESNET_ARM=graph MAGGY_B200_BN=aten timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
RESNET_ARM=eager timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
# launch list of the graph step with fused norms (kernel nodes of one replay)
RESNET_ITERS=1 timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv -s 3000 -c 1200 \
    python scripts/bench_resnet_step.py > $OUT/prof.log 2>&1
python scripts/launch_shares.py $OUT/launches.csv $OUT/shares.txt | head -42
