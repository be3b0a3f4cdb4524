#!/usr/bin/env python
"""BASELINE.json config 5: LOCO ablation study over a Llama-3 style decoder; every trial trains one ablated
variant data-parallel on a gang of GPUs (default: all) with the fused NVLink path. Reports tokens/s per trial
(device-timed inside train_fn, max over gang ranks) and trials/hour.

    python bench/ablation_llama.py --model llama3-1b --gpus-per-trial 8 --steps 10 --batch 4 --seq 2048
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_fns(model_name, batch, seq, steps, mode="dp", fp8=False):
    def base_model():
        from maggy_b200.models.llama import Llama, LlamaConfig

        cfg = LlamaConfig.named(model_name)
        cfg.fp8 = fp8
        return Llama(cfg)

    def dataset(ablated_feature=None):
        return None  # synthetic tokens are generated on the device inside train_fn

    def train_fn(hparams, reporter):
        import torch
        import torch.distributed as dist

        from maggy_b200.parallel.arena import FlatParams, FusedAdamW

        torch.manual_seed(0)
        dev = torch.device("cuda", torch.cuda.current_device())
        with torch.device(dev):  # initialise the 1.2 B parameters on the GPU, not with CPU randn + H2D copies
            model = hparams["model_function"]()
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        opt_sh = None
        if mode == "fsdp-mx" and world > 1:
            # BASELINE config 5 as specified: ZeRO-3 shards, block-scaled fp8, all-gather fused into the GEMM (K3-MX)
            from maggy_b200.models.llama import shard_llama_mx
            from maggy_b200.parallel.fsdp import ShardedLinearMX

            shard_llama_mx(model)
            shard_ids = {id(p) for m in model.modules() if isinstance(m, ShardedLinearMX) for p in m.parameters()}
            flat = FlatParams(model, device=dev, world_size=world, rank=rank, include=lambda n, p: id(p) not in shard_ids)
            flat_sh = FlatParams(model, device=dev, world_size=1, rank=0, include=lambda n, p: id(p) in shard_ids,
                                 use_symmetric=False)
            opt_sh = FusedAdamW(flat_sh, lr=1e-4, grad_scale=1.0)   # the reduce-scatter kernel already averaged
        else:
            flat = FlatParams(model, device=dev, world_size=world, rank=rank)
        opt = FusedAdamW(flat, lr=1e-4)
        vocab = model.cfg.vocab_size
        g = torch.Generator().manual_seed(rank)
        idx = torch.randint(0, vocab, (batch, seq), generator=g).to(dev)
        tgt = torch.randint(0, vocab, (batch, seq), generator=g).to(dev)
        loss, t_ms = None, 0.0
        for step in range(steps + 2):
            if step == 2:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            opt.zero_grad()
            if opt_sh is not None:
                opt_sh.zero_grad()
            loss = model(idx, tgt)
            loss.backward()
            opt.step()
            if opt_sh is not None:
                opt_sh.step()
            reporter.broadcast(-loss.detach(), step)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tok_s = batch * seq * world * steps / (t.item() / 1e3)
        reporter.log(f"tokens/s {tok_s:.0f}", True)
        return {"Metric": -float(loss), "tokens_per_s": tok_s}

    return base_model, dataset, train_fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-1b")
    ap.add_argument("--gpus-per-trial", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--layers", type=int, default=3, help="how many single-block ablations to include")
    ap.add_argument("--mode", default="dp", choices=["dp", "fsdp-mx"],
                    help="dp: replicated bf16 + fused reduce-scatter/Adam (K1); fsdp-mx: ZeRO-3 shards, block-scaled "
                         "fp8, all-gather fused into the GEMM (K3-MX) — BASELINE config 5 as specified")
    ap.add_argument("--fp8", default="", choices=["", "tensor", "mx"], help="dp mode: fp8 forward of the projections")
    args = ap.parse_args()
    import torch

    from maggy_b200 import experiment
    from maggy_b200.ablation import AblationStudy
    from maggy_b200.config import AblationConfig

    ngpu = torch.cuda.device_count()
    gpt = args.gpus_per_trial or max(1, ngpu)
    seq = min(args.seq, 128) if args.model == "llama-tiny" else args.seq
    base_model, dataset, train_fn = make_fns(args.model, args.batch, seq, args.steps, args.mode, args.fp8 or False)
    study = AblationStudy("synthetic_tokens", 1, label_name="next_token", dataset_generator=dataset)
    study.model.set_base_model_generator(base_model)
    study.model.layers.include(*[f"blocks.{i}" for i in range(1, 1 + args.layers)])
    cfg = AblationConfig(ablation_study=study, ablator="loco", direction="max", name="loco_llama",
                         gpus_per_trial=gpt, num_workers=max(1, ngpu // gpt))
    t0 = time.time()
    res = experiment.lagom(train_fn, cfg)
    wall = time.time() - t0
    from maggy_b200.core.environment.singleton import EnvSing

    env = EnvSing.get_instance()
    toks = []
    app = sorted(os.listdir(env.log_dir))[-1]
    run = os.path.join(env.log_dir, app, sorted(os.listdir(os.path.join(env.log_dir, app)), key=int)[-1])
    for tid in os.listdir(run):
        p = os.path.join(run, tid, ".outputs.json")
        if os.path.exists(p):
            toks.append(json.load(open(p)).get("tokens_per_s"))
    toks = [t for t in toks if t]
    kind = "mxfp8 ZeRO-3 (K3-MX)" if args.mode == "fsdp-mx" else ("bf16" if not args.fp8 else f"fp8-{args.fp8} forward")
    print(json.dumps({"metric": f"LOCO ablation, {args.model} {kind}, {gpt} GPUs/trial: tokens/s per trial and trials/hour",
                      "mode": args.mode,
                      "tokens_per_s_median": sorted(toks)[len(toks) // 2] if toks else None, "tokens_per_s_all": toks,
                      "trials": res["num_trials"], "trials_per_hour": res["num_trials"] / wall * 3600.0, "wall_s": wall,
                      "n_gpus": ngpu, "gpus_per_trial": gpt, "best_config": res["best_config"],
                      "config": {"batch_per_gpu": args.batch, "seq": seq, "steps": args.steps}}))


if __name__ == "__main__":
    main()
