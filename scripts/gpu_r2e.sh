#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "norm or bias_gradient or gpt2 or layer_norm or bert" > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/tests.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee -a $OUT/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"; }
run MAGGY_B200_NORM_BWD=2 MAGGY_B200_LN_COL=0
run MAGGY_B200_NORM_BWD=4 MAGGY_B200_LN_COL=0
run MAGGY_B200_NORM_BWD=4 MAGGY_B200_LN_COL=1
run MAGGY_B200_NORM_BWD=2 MAGGY_B200_LN_COL=0
run MAGGY_B200_NORM_BWD=4 MAGGY_B200_LN_COL=1
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
