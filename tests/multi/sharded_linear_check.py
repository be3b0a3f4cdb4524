"""Run under torchrun (>= 2 ranks): ShardedLinear (ZeRO-3 weights gathered inside the forward GEMM, K3) trained for a
few steps against an unsharded reference with all-reduced gradients. Prints SHARDED_LINEAR_OK."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


class MLP(torch.nn.Module):
    def __init__(self, d, hidden):
        super().__init__()
        self.fc1 = torch.nn.Linear(d, hidden)
        self.fc2 = torch.nn.Linear(hidden, d, bias=False)

    def forward(self, x):
        return self.fc2(torch.nn.functional.gelu(self.fc1(x), approximate="tanh"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from maggy_b200.parallel.fsdp import ShardedLinear, shard_linears

    d, hidden = 256 * world, 512 * world
    torch.manual_seed(11)
    ref = MLP(d, hidden).to(dev).to(torch.bfloat16)           # identical on every rank
    model = MLP(d, hidden).to(dev).to(torch.bfloat16)
    model.load_state_dict(ref.state_dict())
    n = shard_linears(model, min_out_features=256)
    assert n == 2 and isinstance(model.fc1, ShardedLinear), n
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.05)
    torch.manual_seed(100 + rank)
    for step in range(4):
        x = torch.randn(512, d, device=dev).to(torch.bfloat16)   # different per rank
        y = torch.randn(512, d, device=dev).to(torch.bfloat16)
        loss = torch.nn.functional.mse_loss(model(x).float(), y.float())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        loss_ref = torch.nn.functional.mse_loss(ref(x).float(), y.float())
        opt_ref.zero_grad(set_to_none=True)
        loss_ref.backward()
        for p in ref.parameters():
            dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
        opt_ref.step()
        assert abs(loss.item() - loss_ref.item()) < 3e-2 * max(1.0, abs(loss_ref.item())), (step, loss.item(), loss_ref.item())
        torch.cuda.synchronize()
        dist.barrier()
    # the updated shards must equal the reference rows
    per = hidden // world
    mine = model.fc1.weight.float()
    want = ref.fc1.weight[rank * per:(rank + 1) * per].float()
    err = ((mine - want).norm() / want.norm()).item()
    assert err < 2e-2, err
    assert model.fc1.weight.data_ptr() == model.fc1.sharded.shard.data_ptr()
    dist.barrier()
    if rank == 0:
        print("SHARDED_LINEAR_OK", f"weight_rel_err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
