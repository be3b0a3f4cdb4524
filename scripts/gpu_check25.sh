#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
  MAGGY_B200_LN_RES=$v timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_lnres_$v.log 2>&1
  echo "LN_RES=$v $(grep '^{' gpurun_out/bench_lnres_$v.log | tail -1 | cut -c98-160)"
done
