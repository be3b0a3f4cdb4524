#!/bin/bash
OUT=gpurun_out/r2u
mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py -m gpu -q -x > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/tests.log
for v in 0 1 0 1; do echo "== DGELU_COLSUM=$v"; MAGGY_B200_DGELU_COLSUM=$v timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['gpu_launches'], d['final_loss'])"; done
