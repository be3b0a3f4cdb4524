#!/bin/bash
# attention pass H: bisect the forward scheduling variants (same box, same clocks)
OUT=gpurun_out/attn_r2h
mkdir -p $OUT
J=$OUT/bench.jsonl
rm -f $J
for rep in 1 2; do
for v in 0 1 2 3; do
  MAGGY_B200_ATTN_VAR=$v timeout 300 python scripts/bench_attn.py --tag var$v --no-ref --shapes 64x1024x12,8x4096x16 --json $J 2>&1 | tee -a $OUT/bench.log
done
MAGGY_B200_ATTN_FWD_ORDER=1 timeout 300 python scripts/bench_attn.py --tag var0_headmajor --no-ref --shapes 64x1024x12,8x4096x16 --json $J 2>&1 | tee -a $OUT/bench.log
done
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -3
