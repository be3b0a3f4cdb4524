"""Run under torchrun with >= 2 ranks: fused reduce-scatter + AdamW + all-gather (K1/K2) vs an NCCL
all-reduce + fp32 AdamW oracle, on the real GPT-2 parameter arena. Prints RS_ADAM_OK on success."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from maggy_b200.models.gpt2 import GPT2, GPT2Config
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    cfg = GPT2Config.named("gpt2-tiny")
    for use_mc in (False, True):
        torch.manual_seed(0)
        model = GPT2(cfg)
        flat = FlatParams(model, device=dev, world_size=world, rank=rank)
        opt = FusedAdamW(flat, lr=1e-2, weight_decay=0.1, use_multicast=use_mc)
        if use_mc and not opt.uses_multicast:
            if rank == 0:
                print("multicast not available on this box: skipping the NVLS variant")
            continue
        # oracle state: full fp32 master + Adam moments on every rank
        ref_p = flat.params.float().clone()
        ref_m, ref_v = torch.zeros_like(ref_p), torch.zeros_like(ref_p)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        for step in range(1, 4):
            idx = torch.randint(0, cfg.vocab_size, (2, 64), generator=g).to(dev)
            tgt = torch.randint(0, cfg.vocab_size, (2, 64), generator=g).to(dev)
            opt.zero_grad()
            loss = model(idx, tgt)
            loss.backward()
            flat.zero_unwritten_grads()
            gsum = flat.grads.float().clone()
            dist.all_reduce(gsum)
            gavg = gsum / world
            opt.step()
            torch.cuda.synchronize()
            # oracle AdamW (decoupled decay on [0, decay_boundary))
            b1, b2, eps, lr = 0.9, 0.95, 1e-8, 1e-2
            ref_m.mul_(b1).add_(gavg, alpha=1 - b1)
            ref_v.mul_(b2).addcmul_(gavg, gavg, value=1 - b2)
            mhat = ref_m / (1 - b1 ** step)
            denom = (ref_v / (1 - b2 ** step)).sqrt() + eps
            wd = torch.zeros_like(ref_p)
            wd[:flat.decay_boundary] = 0.1
            ref_p = ref_p - lr * (mhat / denom + wd * ref_p)
            got = flat.params.float()
            # bf16 gradients summed in different order + bf16 parameter rounding
            err = (got - ref_p).abs().max().item()
            scale = ref_p.abs().max().item()
            assert err < 2e-2 * max(scale, 1.0), (use_mc, step, err, scale)
            rel = ((got - ref_p).norm() / ref_p.norm()).item()
            assert rel < 5e-3, (use_mc, step, rel)
            # every rank must hold bit-identical parameters after the all-gather
            mine = flat.params.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(mine, other), f"rank {rank} parameters differ from rank 0 (mc={use_mc}, step {step})"
        if rank == 0:
            print(f"fused rs+adam+ag ok (multicast={use_mc and opt.uses_multicast}) loss={loss.item():.4f}")
        dist.barrier()

    # CUDA-graph replay: device-side epoch/step counters must keep the ranks in lock-step
    # (MAGGY_CHECK_GRAPH=0: eager, for compute-sanitizer runs — stream capture fails under the tool)
    from maggy_b200.parallel.trainer import build_gpt2_trainer

    use_graph = os.environ.get("MAGGY_CHECK_GRAPH", "1") != "0"
    model, flat, opt, step = build_gpt2_trainer(cfg, 2, 64, dev, world_size=world, rank=rank, use_cuda_graph=use_graph)
    idx = torch.randint(0, cfg.vocab_size, (2, 64)).pin_memory()
    tgt = torch.randint(0, cfg.vocab_size, (2, 64)).pin_memory()
    losses = [step(idx, tgt).item() for _ in range(8)]
    assert losses[-1] < losses[0], losses
    mine = flat.params.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other)
    dist.barrier()
    if rank == 0:
        print("RS_ADAM_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
