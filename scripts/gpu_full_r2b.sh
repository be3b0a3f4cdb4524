#!/bin/bash
# full GPU tier (what the driver runs at round end), time-boxed
OUT=gpurun_out/full_r2b
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/gputests.log 2>&1
tail -6 $OUT/gputests.log
grep -E "^E " $OUT/gputests.log | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
