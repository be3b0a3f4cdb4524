"""ZeRO-3 style linear layers whose weight all-gather happens INSIDE the forward GEMM (K3).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/zero3_sharded_linear.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from maggy_b200.parallel.fsdp import ShardedLinear, shard_linears


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    torch.manual_seed(0)                                             # same initial weights on every rank
    model = torch.nn.Sequential(torch.nn.Linear(1024, 2048 * world), torch.nn.GELU(approximate="tanh"),
                                torch.nn.Linear(2048 * world, 1024)).to(dev).to(torch.bfloat16)
    n = shard_linears(model)                                         # each rank now owns 1/world of the rows
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    torch.manual_seed(100 + rank)                                    # different data per rank
    for step in range(5):
        x = torch.randn(4096, 1024, device=dev).to(torch.bfloat16)
        loss = model(x).float().pow(2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()                                              # wgrad -> reduce-scatter: shard-sized gradients
        opt.step()                                                   # in place on the symmetric-memory shard
        if rank == 0:
            print(f"step {step}: loss {loss.item():.4f} ({n} layers sharded, "
                  f"{sum(isinstance(m, ShardedLinear) for m in model.modules())} ShardedLinear modules)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
