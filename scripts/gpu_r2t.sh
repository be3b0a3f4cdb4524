#!/bin/bash
OUT=gpurun_out/r2t
mkdir -p $OUT
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gputests.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.log 2>&1; grep '^{' $OUT/bench_n1.log | tail -1 > $OUT/bench_n1.json; cut -c1-210 $OUT/bench_n1.json
RESNET_ARM=graph timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee $OUT/step.log
