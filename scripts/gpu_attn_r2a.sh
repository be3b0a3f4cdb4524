#!/bin/bash
# round-2 attention pass A: numerics of generation 2, variant sweep, ncu of fwd + bwd
set -x
mkdir -p gpurun_out
OUT=gpurun_out/attn_r2a
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -15 > $OUT/tests.log
cat $OUT/tests.log
J=$OUT/bench.jsonl
rm -f $J
timeout 300 python scripts/bench_attn.py --tag gen2 --json $J 2>&1 | tee $OUT/bench_gen2.log
MAGGY_B200_ATTN_GEN=1 timeout 300 python scripts/bench_attn.py --tag gen1 --no-ref --shapes 64x1024x12 --json $J 2>&1 | tee $OUT/bench_gen1.log
MAGGY_B200_ATTN_POLY=1 MAGGY_B200_ATTN_POLY_BWD=1 timeout 300 python scripts/bench_attn.py --tag poly1 --no-ref --shapes 64x1024x12,8x4096x16 --json $J 2>&1 | tee $OUT/bench_poly1.log
MAGGY_B200_ATTN_POLY=2 MAGGY_B200_ATTN_BWD_ORDER=0 timeout 300 python scripts/bench_attn.py --tag poly2_order0 --no-ref --shapes 64x1024x12,8x4096x16 --json $J 2>&1 | tee $OUT/bench_poly2.log
kill $SMI
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_v2 -s 4 -c 1 -o $OUT/prof_attn_fwd_v2 python scripts/prof_attn.py > $OUT/ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_v2 -s 1 -c 1 -o $OUT/prof_attn_bwd_v2 python scripts/prof_attn.py > $OUT/ncu_bwd.log 2>&1
ls -la $OUT
