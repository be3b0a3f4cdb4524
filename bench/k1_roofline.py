#!/usr/bin/env python
"""Roofline of the fused reduce-scatter + AdamW + all-gather kernel (K1/K2) on the GPT-2 small arena.

Run under torchrun with N >= 2 ranks. Device-timed (CUDA events around the single kernel launch, barrier +
synchronize on both sides, max over ranks), compared with
  (a) its roofline = max(bytes over NVLink / 770 GB/s per direction, HBM bytes / measured copy bandwidth),
  (b) the library path it replaces: NCCL reduce_scatter + torch fused AdamW on the shard + NCCL all_gather.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from maggy_b200.models.gpt2 import GPT2, GPT2Config
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    torch.manual_seed(0)
    model = GPT2(GPT2Config.named("gpt2-small"))
    flat = FlatParams(model, device=dev, world_size=world, rank=rank)
    n = flat.numel
    results = {}
    for use_mc in (True, False):
        opt = FusedAdamW(flat, lr=1e-4, use_multicast=use_mc)
        if use_mc and not opt.uses_multicast:
            continue
        flat.grads.normal_(0, 1e-3)
        for p in flat.parameters():
            p._mg_written = True
        ts = []
        for it in range(13):
            for p in flat.parameters():
                p._mg_written = True
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.step()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 3:
                ts.append(t.item())
        ts.sort()
        results["fused_mc" if use_mc else "fused_p2p"] = ts[len(ts) // 2]

    # library path: reduce_scatter (bf16) + fused AdamW on the fp32 shard + cast + all_gather (bf16)
    shard = n // world
    g_shard = torch.empty(shard, dtype=torch.bfloat16, device=dev)
    master = flat.master.clone().requires_grad_(True)
    topt = torch.optim.AdamW([master], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.1, fused=True)
    pb = torch.empty(shard, dtype=torch.bfloat16, device=dev)
    full = torch.empty(n, dtype=torch.bfloat16, device=dev)
    grads = flat.grads.clone()
    ts = []
    for it in range(13):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce_scatter_tensor(g_shard, grads)
        master.grad = g_shard.float() / world
        topt.step()
        pb.copy_(master.detach())
        dist.all_gather_into_tensor(full, pb)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 3:
            ts.append(t.item())
    ts.sort()
    results["nccl_rs_adam_ag"] = ts[len(ts) // 2]

    peaks = {"hbm_gbs": 6576.4}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    link = 770e9  # measured peer copy bandwidth per direction (profiling guide)
    nvlink_bytes = (world - 1) / world * n * 2          # per direction per GPU (grads in, params out)
    hbm_bytes = shard * (24 + 2 + 2) + n * 2 / world    # fp32 master/m/v read+write, bf16 grad read, bf16 param write
    t_link, t_hbm = nvlink_bytes / link * 1e3, hbm_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
    roof = max(t_link, t_hbm)
    best = min(v for k, v in results.items() if k.startswith("fused"))
    out = {"kernel": "rs_adam (reduce-scatter + AdamW + all-gather)", "n_gpus": world, "params": n,
           "ms": results, "roofline_ms": {"nvlink": t_link, "hbm": t_hbm, "bound": roof},
           "fraction_of_roofline": roof / best, "speedup_vs_nccl_path": results["nccl_rs_adam_ag"] / best,
           "nvlink_bytes_per_dir": nvlink_bytes, "hbm_bytes": hbm_bytes}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
