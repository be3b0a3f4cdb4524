"""Data-parallel module wrapper on the fused NVLink path (replaces torch DDP + ZeRO for modules built
on maggy_b200 ops; reference: maggy/core/patching/modules.py:38-65).

``DataParallel(module)`` re-homes the parameters into symmetric-memory arenas (``FlatParams``); the
gradient all-reduce DDP would do in bucket hooks is instead *fused into the optimizer step*:
``FusedAdamW.step()`` = reduce-scatter + Adam + all-gather in one kernel (csrc/comm/rs_adam.cu).
Rank 0's initial parameters are the ones every rank starts from (DDP's constructor broadcast,
reference SURVEY 2.4 C3): all ranks seed identically, and the arena copy is verified with a checksum.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .arena import FlatParams, FusedAdamW


class DataParallel(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, device=None, group=None, broadcast_params: bool = True):
        super().__init__()
        self.module = module
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.flat = FlatParams(module, device=device, world_size=world, rank=rank, group=group)
        if world > 1 and broadcast_params:
            # make rank 0 authoritative (cheap: one bf16 arena), then rebuild the fp32 master shard from it
            dist.broadcast(self.flat.params, src=0, group=group)
            lo = rank * self.flat.shard_numel
            self.flat.master.copy_(self.flat.params[lo:lo + self.flat.shard_numel].float())
        self._optimizer: Optional[FusedAdamW] = None

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def make_optimizer(self, **kw) -> FusedAdamW:
        """The fused sharded AdamW bound to this module's arenas."""
        self._optimizer = FusedAdamW(self.flat, **kw)
        return self._optimizer

    def parameters(self, recurse: bool = True):
        return self.module.parameters(recurse)

    def named_parameters(self, *a, **k):
        return self.module.named_parameters(*a, **k)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)
