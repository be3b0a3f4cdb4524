#!/bin/bash
OUT=gpurun_out/diag
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi/generic_dp_check.py > $OUT/generic.out 2> $OUT/generic.err
timeout 200 python -m pytest tests/test_gpu_lagom.py -x -q -k device_scheduler > $OUT/lagom.out 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "parity" > $OUT/parity.out 2>&1
grep -E "Error|error:|assert" $OUT/generic.err | head -10
