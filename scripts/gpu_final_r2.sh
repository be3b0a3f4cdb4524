#!/bin/bash
# round-2 validation on one GPU: the full GPU test tier (as the driver runs it), smoke(), the headline bench, the ResNet step
OUT=gpurun_out/final_r2
mkdir -p $OUT
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -5 $OUT/gputests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.log 2>&1; grep '^{' $OUT/bench_n1.log | tail -1 > $OUT/bench_n1.json; cat $OUT/bench_n1.json | cut -c1-400
RESNET_ARM=graph timeout 300 python scripts/bench_resnet_step.py 2>&1 | tail -1 | tee $OUT/resnet_step.json
BN_STEP=0 timeout 300 python scripts/bench_bn.py 2>&1 | tail -8 | tee $OUT/bench_bn.log
BW_ONLY_NORM=1 timeout 120 python scripts/bench_bw_kernels.py 2>&1 | tail -4 | tee $OUT/bw_norm.log
