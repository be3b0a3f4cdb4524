#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"attn_" -s 2 -c 3 \
   -o gpurun_out/prof_attn python scripts/prof_attn.py > gpurun_out/prof_attn.log 2>&1
tail -3 gpurun_out/prof_attn.log
