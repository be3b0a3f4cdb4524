#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t_full.log
tail -8 gpurun_out/t_full.log
timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final1.log 2>&1
grep '^{' gpurun_out/bench_final1.log | tail -1
python -c "
import __graft_entry__ as g, time
t=time.time(); g.smoke(); print('smoke ok', round(time.time()-t,1),'s')" 2>&1 | tail -2
