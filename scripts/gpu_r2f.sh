#!/bin/bash
OUT=gpurun_out/r2f
mkdir -p $OUT
for impl in native torch; do
  PROF_IMPL=$impl timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_$impl.csv \
    python scripts/prof_resnet_step.py > $OUT/prof_$impl.log 2>&1
  echo "ncu $impl rc=$?"; tail -1 $OUT/prof_$impl.log
  python scripts/launch_shares.py $OUT/launches_$impl.csv $OUT/shares_$impl.txt | head -45
done
