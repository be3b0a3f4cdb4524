#!/bin/bash
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/gputests.log
timeout 300 python scripts/list_smoke_kernels.py > $OUT/smoke_kernels.log 2>&1; tail -25 $OUT/smoke_kernels.log
