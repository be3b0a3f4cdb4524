#!/bin/bash
# last GPU call of round 2: full test tier on the final tree + launch lists of both steps in their final state
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gputests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
RESNET_ITERS=1 timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_resnet.csv -s 3000 -c 900 \
    python scripts/bench_resnet_step.py > $OUT/prof_resnet.log 2>&1
python scripts/launch_shares.py $OUT/launches_resnet.csv $OUT/shares_resnet.txt | head -12
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_gpt2.csv \
  python scripts/prof_step.py > $OUT/prof_gpt2.log 2>&1
python scripts/launch_shares.py $OUT/launches_gpt2.csv $OUT/shares_gpt2.txt | head -16
