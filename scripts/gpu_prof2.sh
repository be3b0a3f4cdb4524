#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python bench.py --steps 20 --warmup 5 --batch 64 > gpurun_out/bench_n1_b64.log 2>&1
timeout -k 10 200 python bench.py --steps 20 --warmup 5 --batch 32 > gpurun_out/bench_n1_b32.log 2>&1
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 520 --csv \
   --log-file gpurun_out/launches_v2.csv python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/launches_v2.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 3 -c 2 \
   -o gpurun_out/prof_gemm_v2 python scripts/prof_gemm.py > gpurun_out/prof_gemm_v2.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"adam_flat|ce_fwd_bwd|norm_fwd|norm_bwd_kernel" -s 40 -c 8 \
   -o gpurun_out/prof_bw_kernels python bench.py --steps 1 --warmup 3 --no-graph --batch 32 > gpurun_out/prof_bw.log 2>&1
ls -la gpurun_out | tail -8
for f in bench_n1_b64 bench_n1_b32; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-260; done
