#!/bin/bash
# 8-GPU lease, end of round 2: headline bench at N=8 through lagom, ASHA trials/hour (BASELINE config 2) with the host
# engine (8 workers) and with the device scheduler on the spare GPU (7 workers + K9), and the stock-kernel eager arm
OUT=gpurun_out/n8_r2b
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "bench n8 rc=$?"; grep '^{' $OUT/bench_n8.json | tail -1 | cut -c1-600; tail -2 $OUT/bench_n8.err
timeout 400 python bench.py --config asha --gpus 8 --trials 64 --steps-per-budget 200 > $OUT/asha_n8.json 2> $OUT/asha_n8.err
echo "asha n8 rc=$?"; tail -c 1400 $OUT/asha_n8.json; tail -2 $OUT/asha_n8.err
timeout 400 python bench/hpo_asha_resnet.py --workers 7 --trials 64 --steps-per-budget 200 > $OUT/asha_n7_k9.json 2> $OUT/asha_n7_k9.err
echo "asha n7+k9 rc=$?"; tail -c 1600 $OUT/asha_n7_k9.json; tail -2 $OUT/asha_n7_k9.err
timeout 300 python bench/hpo_asha_resnet.py --workers 8 --trials 32 --steps-per-budget 100 --engine eager > $OUT/asha_n8_eager.json 2> $OUT/asha_n8_eager.err
echo "asha eager rc=$?"; tail -c 900 $OUT/asha_n8_eager.json; tail -2 $OUT/asha_n8_eager.err
timeout 300 python bench/hpo_asha_resnet.py --workers 8 --trials 32 --steps-per-budget 100 > $OUT/asha_n8_graph_32x100.json 2> $OUT/asha_n8_graph_32x100.err
echo "asha graph 32x100 rc=$?"; tail -c 900 $OUT/asha_n8_graph_32x100.json
