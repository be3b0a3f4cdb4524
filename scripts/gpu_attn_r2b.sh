#!/bin/bash
# round-2 attention pass B: pair-ordered forward MMA issue, 16-warp backward compute, slim waits
set -x
OUT=gpurun_out/attn_r2b
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -15 > $OUT/tests.log
cat $OUT/tests.log
J=$OUT/bench.jsonl
rm -f $J
timeout 300 python scripts/bench_attn.py --tag gen2b --json $J 2>&1 | tee $OUT/bench_gen2b.log
MAGGY_B200_ATTN_POLY=1 timeout 300 python scripts/bench_attn.py --tag poly1_fwd --no-ref --shapes 64x1024x12 --json $J 2>&1 | tee $OUT/bench_poly1.log
MAGGY_B200_ATTN_BWD_ORDER=0 timeout 300 python scripts/bench_attn.py --tag order0 --no-ref --shapes 64x1024x12 --json $J 2>&1 | tee $OUT/bench_order0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_v2 -s 4 -c 1 -o $OUT/prof_attn_fwd_v2 python scripts/prof_attn.py > $OUT/ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_v2 -s 1 -c 1 -o $OUT/prof_attn_bwd_v2 python scripts/prof_attn.py > $OUT/ncu_bwd.log 2>&1
ls -la $OUT
