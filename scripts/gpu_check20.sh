#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python scripts/bench_bw_kernels.py gpurun_out/bw_kernels.json > gpurun_out/bw_kernels.log 2>&1
tail -14 gpurun_out/bw_kernels.log
