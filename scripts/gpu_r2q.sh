#!/bin/bash
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 400 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $OUT/ref_n1.log 2>&1; echo "reference arm rc=$?"; grep '^{' $OUT/ref_n1.log | tail -1 | cut -c1-500
timeout 300 python bench.py > $OUT/ours_default.log 2>&1; echo "default bench rc=$?"; grep '^{' $OUT/ours_default.log | tail -1 | cut -c1-300
