#!/bin/bash
# compute-sanitizer passes, round 2: memcheck over the new kernels (attention gen 2, MXFP8 GEMM + quantiser,
# device mailbox), racecheck + synccheck over GEMM / attention / MXFP8, memcheck over a 2-rank K1 + K3-MX + rs_sum run.
# Every step is bounded: the tools slow kernels down by 10-100x.
OUT=gpurun_out/sanitizer_r2
mkdir -p $OUT
SEL="(attention_fwd and case0) or (attention_bwd and case0) or (gemm_mxfp8 and mnk0) or (quantize_mxfp8 and shape0) or report_store_publishes"
timeout -k 10 240 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_attention.py tests/test_gpu_fp8.py tests/test_gpu_ctrl.py -m gpu -q -x -k "$SEL" > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck.log | tail -3
SEL2="(attention_fwd and case0) or (gemm_mxfp8 and mnk0-False) or (gemm_epilogues)"
timeout -k 10 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_attention.py tests/test_gpu_fp8.py tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL2" > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" $OUT/racecheck.log | tail -3
timeout -k 10 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_attention.py tests/test_gpu_fp8.py tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL2" > $OUT/synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/synccheck.log | tail -3
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout -k 10 400 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multi/rs_adam_check.py > $OUT/memcheck_k1_n2.log 2>&1
  echo "memcheck K1 n2 rc=$?"; grep -E "ERROR SUMMARY|RS_ADAM_OK" $OUT/memcheck_k1_n2.log | tail -4
fi
# launch list of one training step (shares, not absolutes): which kernels the step consists of after this round
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-graph > $OUT/launch_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/sanitizer_r2/launches.csv") if not l.startswith("==")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError: continue
    name = r[ix["Kernel Name"]].split("(")[0][-60:]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
with open("gpurun_out/sanitizer_r2/launch_shares.txt", "w") as f:
    for k, v in tot.most_common(28):
        line = f"{100*v/s:5.1f}%  x{cnt[k]:4d}  {k}"
        print(line); f.write(line + "\n")
PY
