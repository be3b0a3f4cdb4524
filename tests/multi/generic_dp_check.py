"""Run under torchrun (>= 2 ranks): a plain nn.Sequential MLP (fp32, torch autograd, nothing from maggy_b200.models)
under the class-factory wrapper + scoped optimizer patch trains through the fused fp32-arena reduce-scatter/Adam/all-gather
kernel and matches the DDP + AdamW oracle. Also checks the K2 peer broadcast. Prints GENERIC_DP_OK."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


class MLP(torch.nn.Module):
    def __init__(self, d=96, h=200, out=10):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(d, h), torch.nn.ReLU(), torch.nn.LayerNorm(h),
                                       torch.nn.Linear(h, out))

    def forward(self, x):
        return self.net(x)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from maggy_b200.core.patching.modules import get_maggy_ddp_wrapper
    from maggy_b200.core.patching.optim import patched_optimizers
    from maggy_b200.parallel.arena import FusedAdamW

    for zero_lvl, opt_name in [(1, "AdamW"), (0, "AdamW"), (1, "SGD")]:
        torch.manual_seed(1234 + rank)        # DIFFERENT initial weights per rank: the wrapper must broadcast rank 0's
        Wrapped = get_maggy_ddp_wrapper(MLP)
        model = Wrapped()
        assert model.fused, "generic fp32 module did not take the fused path"
        torch.manual_seed(1234)               # oracle = rank 0's initialisation (built on the device, like the wrapper)
        with torch.device(dev):
            ref = MLP()
        for (n1, p1), (n2, p2) in zip(model.inner.module.named_parameters(), ref.named_parameters()):
            assert torch.equal(p1.detach(), p2.detach()), f"K2 broadcast: {n1} differs from rank 0's value"
        ref = torch.nn.parallel.DistributedDataParallel(ref, device_ids=[dev.index])
        with patched_optimizers(zero_lvl):
            if opt_name == "AdamW":
                opt = torch.optim.AdamW(model.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0)
                assert isinstance(opt, FusedAdamW)
            else:
                opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0) if opt_name == "AdamW" \
            else torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
        for step in range(6):
            torch.manual_seed(99 + step * world + rank)     # different data per rank
            x = torch.randn(32, 96, device=dev)
            y = torch.randint(0, 10, (32,), device=dev)
            opt.zero_grad()
            torch.nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
            ropt.zero_grad()
            torch.nn.functional.cross_entropy(ref(x), y).backward()
            ropt.step()
        for (n1, p1), (n2, p2) in zip(model.inner.module.named_parameters(), ref.module.named_parameters()):
            err = ((p1.detach() - p2.detach()).norm() / (p2.detach().norm() + 1e-12)).item()
            assert err < 2e-3, (zero_lvl, opt_name, n1, err)
        # every rank holds the same parameters, bit for bit
        flat = model.inner.flat
        lo, hi = flat.params.clone(), flat.params.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), (zero_lvl, opt_name, "ranks diverged")
    if rank == 0:
        print("GENERIC_DP_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
