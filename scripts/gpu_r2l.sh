#!/bin/bash
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv.py tests/test_gpu_batchnorm.py tests/test_gpu_fp8.py -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/tests.log
for pdl in 0 1; do
  echo "== PDL=$pdl"
  MAGGY_B200_PDL=$pdl RESNET_ARM=graph timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee -a $OUT/step.log
  MAGGY_B200_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee -a $OUT/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
done
