#!/usr/bin/env python
"""BASELINE.json config 4: Bayesian optimization (GP) + median early stopping over BERT-base masked-LM
pre-training on synthetic tokens, every trial on a gang of GPUs (default 2) running the fused NVLink
data-parallel path (maggy_b200.parallel.DataParallel + FusedAdamW) inside the oblivious train_fn.

    python bench/hpo_gp_bert.py --gpus-per-trial 2 --trials 12 --steps 30 --batch 16
    python bench/hpo_gp_bert.py --cpu --model bert-tiny --trials 6 --gpus-per-trial 1       # plumbing check
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_train_fn(model_name, batch, seq, steps, cpu):
    def train_fn(hparams, reporter):
        import torch
        import torch.distributed as dist

        from maggy_b200.models.bert import BertConfig, BertForMaskedLM

        cfg = BertConfig.named(model_name)
        torch.manual_seed(0)
        if cpu:  # plumbing mode: a stand-in objective, no kernels
            loss = 0.0
            for step in range(steps):
                loss = (hparams["lr"] - 3e-4) ** 2 * 1e6 + hparams["wd"] + 1.0 / (step + 1)
                reporter.broadcast(-loss, step)
            return -loss
        from maggy_b200.parallel.data_parallel import DataParallel

        dev = torch.device("cuda", torch.cuda.current_device())
        with torch.device(dev):  # parameters are initialised on the GPU
            model = DataParallel(BertForMaskedLM(cfg), device=dev) if dist.is_initialized() else None
        if model is None:
            from maggy_b200.parallel.arena import FlatParams, FusedAdamW

            with torch.device(dev):
                net = BertForMaskedLM(cfg)
            flat = FlatParams(net, device=dev)
            opt = FusedAdamW(flat, lr=hparams["lr"], weight_decay=hparams["wd"])
            fwd = net
        else:
            opt = model.make_optimizer(lr=hparams["lr"], weight_decay=hparams["wd"])
            fwd = model
        g = torch.Generator(device="cpu").manual_seed(dist.get_rank() if dist.is_initialized() else 0)
        idx = torch.randint(0, cfg.vocab_size, (batch, seq), generator=g).to(dev)
        labels = torch.randint(0, cfg.vocab_size, (batch, seq), generator=g).to(dev)
        loss = None
        for step in range(steps):
            opt.zero_grad()
            loss = fwd(idx, labels)
            loss.backward()
            opt.step()
            if step % 5 == 4 or step == steps - 1:
                reporter.broadcast(-loss.detach(), step)   # device tensor: published by the K8 kernel
        return -float(loss)

    return train_fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus-per-trial", type=int, default=2)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--trials", type=int, default=12)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--model", default="bert-base")
    ap.add_argument("--optimizer", default="gp", choices=["gp", "tpe"])
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    if args.cpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import torch

    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig
    from maggy_b200.experiment import experiment as exp_mod
    from maggy_b200.optimizer.bayes import GP, TPE

    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    workers = args.workers or max(1, (ngpu // args.gpus_per_trial) if ngpu else 2)
    seq = min(args.seq, 128) if args.model == "bert-tiny" else args.seq
    sp = Searchspace(lr=("DOUBLE", [5e-5, 1e-3]), wd=("DOUBLE", [0.0, 0.1]))
    opt = GP(num_warmup_trials=4, acq_optimizer_kwargs={"n_points": 2000}) if args.optimizer == "gp" \
        else TPE(num_warmup_trials=4)
    cfg = HyperparameterOptConfig(num_trials=args.trials, optimizer=opt, searchspace=sp, direction="max",
                                  es_policy="median", es_interval=1, es_min=3, name="gp_bert",
                                  gpus_per_trial=args.gpus_per_trial, num_workers=workers, seed=0)
    holder = {}
    orig = exp_mod.lagom_driver

    def spy(config, app_id, run_id):
        holder["d"] = orig(config, app_id, run_id)
        return holder["d"]

    exp_mod.lagom_driver = spy
    t0 = time.time()
    try:
        res = experiment.lagom(make_train_fn(args.model, args.batch, seq, args.steps, args.cpu), cfg)
    finally:
        exp_mod.lagom_driver = orig
    wall = time.time() - t0
    d = holder["d"]
    gaps = sorted(d.inter_trial_gaps)
    samp = sorted(t.info_dict.get("sampling_time", 0.0) for t in d._final_store)
    print(json.dumps({
        "metric": f"trials/hour ({args.optimizer.upper()} + median early stop, {args.model} MLM synthetic, "
                  f"{args.gpus_per_trial} GPUs/trial)",
        "value": res["num_trials"] / wall * 3600.0, "unit": "trials/hour", "n_gpus": ngpu, "workers": workers,
        "gpus_per_trial": args.gpus_per_trial, "trials": res["num_trials"], "early_stopped": res["early_stopped"],
        "wall_s": wall, "best_val": res["best_val"],
        "inter_trial_gap_ms_median": gaps[len(gaps) // 2] * 1e3 if gaps else None,
        "inter_trial_gap_ms_max": gaps[-1] * 1e3 if gaps else None,
        "suggestion_time_ms_median": samp[len(samp) // 2] * 1e3 if samp else None,
        "suggestion_time_ms_max": samp[-1] * 1e3 if samp else None,
        "note": "inter-trial gap < suggestion time means the GP fit was hidden behind running trials (lookahead)",
    }))


if __name__ == "__main__":
    main()
