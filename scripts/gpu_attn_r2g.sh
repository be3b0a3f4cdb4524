#!/bin/bash
# round-2 attention pass G: final forward micro-fixes; end-to-end effect of native vs sdpa in bench.py
set -x
OUT=gpurun_out/attn_r2g
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -5 > $OUT/tests.log
cat $OUT/tests.log
J=$OUT/bench.jsonl
rm -f $J
timeout 300 python scripts/bench_attn.py --tag gen2g --json $J 2>&1 | tee $OUT/bench_gen2g.log
timeout 300 python bench.py --steps 10 --warmup 3 --attn native > $OUT/bench_native.json 2> $OUT/bench_native.err
timeout 300 python bench.py --steps 10 --warmup 3 --attn sdpa > $OUT/bench_sdpa.json 2> $OUT/bench_sdpa.err
tail -c 1500 $OUT/bench_native.json; tail -c 1500 $OUT/bench_sdpa.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_v2 -s 4 -c 1 -o $OUT/prof_attn_fwd_v2 python scripts/prof_attn.py > $OUT/ncu_fwd.log 2>&1
ls -la $OUT
