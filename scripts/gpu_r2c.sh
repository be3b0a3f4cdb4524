#!/bin/bash
OUT=gpurun_out/r2c
mkdir -p $OUT
for chk in rs_adam_check ag_gemm_check mx_fsdp_check; do
  MAGGY_CHECK_GRAPH=0 timeout -k 10 300 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multi/$chk.py > $OUT/memcheck_$chk.log 2>&1
  echo "memcheck $chk n2 rc=$?"; grep -E "ERROR SUMMARY|_OK" $OUT/memcheck_$chk.log | tail -4
done
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv \
  python scripts/prof_step.py > $OUT/prof_step.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/prof_step.log
python scripts/launch_shares.py $OUT/launches.csv $OUT/launch_shares.txt
