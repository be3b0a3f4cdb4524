#!/bin/bash
mkdir -p gpurun_out
cd /tmp && rm -rf hpo_run && mkdir hpo_run && cd hpo_run
timeout -k 10 300 python $GRAFT_REPO_ROOT/bench/hpo_gp_bert.py --gpus-per-trial 2 --trials 5 --steps 20 --batch 16 > $GRAFT_REPO_ROOT/gpurun_out/hpo_gp_bert_n2.log 2>&1
timeout -k 10 300 python $GRAFT_REPO_ROOT/examples/gpt2_distributed.py > $GRAFT_REPO_ROOT/gpurun_out/ex_gpt2_dist.log 2>&1
cd $GRAFT_REPO_ROOT
tail -2 gpurun_out/hpo_gp_bert_n2.log | cut -c1-1000; tail -4 gpurun_out/ex_gpt2_dist.log | cut -c1-600
for d in /tmp/hpo_run/experiment_log/*/*/; do tail -6 $d/worker_0_0.stdout.log 2>/dev/null | cut -c1-300; done | tail -20
