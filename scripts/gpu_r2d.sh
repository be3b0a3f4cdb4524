#!/bin/bash
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "norm or bias_gradient or gpt2 or layer_norm" > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/tests.log
for v in 2 4; do echo "== NORM_BWD=$v"; MAGGY_B200_NORM_BWD=$v BW_ONLY_NORM=1 timeout 120 python scripts/bench_bw_kernels.py 2>&1 | tee -a $OUT/bw_norm.log | tail -4; done
for col in 0 1; do echo "== LN_COL=$col"; MAGGY_B200_LN_COL=$col timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee -a $OUT/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"; done
