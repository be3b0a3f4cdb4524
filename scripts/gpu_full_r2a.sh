#!/bin/bash
# full GPU test tier + headline bench through lagom (what the driver runs at round end)
OUT=gpurun_out/full_r2a
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -c 2500 $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
