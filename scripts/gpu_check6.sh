#!/bin/bash
mkdir -p gpurun_out
cd /tmp && rm -rf hpo_run && mkdir hpo_run && cd hpo_run
timeout -k 10 300 python $GRAFT_REPO_ROOT/bench/hpo_asha_resnet.py --workers 2 --trials 8 --steps-per-budget 6 --batch 64 > $GRAFT_REPO_ROOT/gpurun_out/hpo_asha_n2.log 2>&1
timeout -k 10 300 python $GRAFT_REPO_ROOT/bench/hpo_asha_resnet.py --workers 2 --trials 8 --steps-per-budget 6 --batch 64 --proxy-reference-rpc > $GRAFT_REPO_ROOT/gpurun_out/hpo_asha_n2_proxy.log 2>&1
timeout -k 10 300 python $GRAFT_REPO_ROOT/bench/hpo_gp_bert.py --gpus-per-trial 2 --trials 5 --steps 20 --batch 16 > $GRAFT_REPO_ROOT/gpurun_out/hpo_gp_bert_n2.log 2>&1
cd $GRAFT_REPO_ROOT
for f in hpo_asha_n2 hpo_asha_n2_proxy hpo_gp_bert_n2; do tail -3 gpurun_out/$f.log | cut -c1-1000; done
ls /tmp/hpo_run/experiment_log/*/ | head; for d in /tmp/hpo_run/experiment_log/*/*/; do tail -5 $d/worker_0_0.stdout.log 2>/dev/null | cut -c1-300; done | tail -30
