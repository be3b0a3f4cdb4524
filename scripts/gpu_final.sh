#!/bin/bash
# final validation of the round: full GPU suite, norm-backward variants (numerics + same-box A/B), bench, smoke
mkdir -p gpurun_out
for v in 2 3; do
  MAGGY_B200_NORM_BWD=$v timeout -k 10 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "norm or gpt2 or llama" 2>&1 | tail -2
done
for v in 1 2 3; do
  MAGGY_B200_NORM_BWD=$v timeout -k 10 100 python scripts/bench_bw_kernels.py 2>/dev/null | grep "layer_norm bwd" | sed "s/^/variant $v: /"
done
for v in 1 2 3 1; do
  MAGGY_B200_NORM_BWD=$v timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_nb_$v.log 2>&1
  echo "NORM_BWD=$v $(grep '^{' gpurun_out/bench_nb_$v.log | tail -1 | cut -c98-160)"
done
timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout -k 10 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.log 2>&1
grep '^{' gpurun_out/bench_final.log | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
