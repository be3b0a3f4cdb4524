#!/bin/bash
mkdir -p gpurun_out
# every launch of one training step with its device time (eager mode so kernels are individually visible)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/launches.log 2>&1
# top kernel, full set
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 3 -c 3 \
   -o gpurun_out/prof_gemm python scripts/prof_gemm.py > gpurun_out/prof_gemm.log 2>&1
ls -la gpurun_out | tail; tail -3 gpurun_out/launches.log; tail -3 gpurun_out/prof_gemm.log
