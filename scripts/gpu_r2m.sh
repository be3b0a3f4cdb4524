#!/bin/bash
OUT=gpurun_out/r2m
mkdir -p $OUT
MAGGY_B200_GRAPH_POOL=private timeout 300 python scripts/prof_trial_startup.py 2>&1 | grep "^trial" | tee $OUT/startup_private.log
MAGGY_B200_GRAPH_POOL=shared timeout 300 python scripts/prof_trial_startup.py 2>&1 | grep "^trial" | tee $OUT/startup_shared.log
