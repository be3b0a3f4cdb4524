"""Whole-step execution engine: forward + backward + fused optimizer as ONE replayable CUDA graph.

B200-first replacement for the per-op eager loop the reference leaves to the user inside train_fn
(maggy/core/executors/torch_dist_executor.py:145-159): the step is launch-bound at small batch on a
148-SM part, so the engine captures it once (CUDA streams + graphs instead of a tracing compiler)
and replays it; per-step host work is one pinned H2D copy of the batch and, optionally, one D2H
read of the loss.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from ..utils import tracing
from .arena import FlatParams, FusedAdamW


_GRAPH_POOLS: dict = {}


def _shared_graph_pool(device):
    """One graph memory pool per device for the life of the process, kept alive by a one-node graph.

    An HPO worker captures a fresh step graph for every trial. With a private pool per graph, destroying trial N's
    graph ``cudaFree``s its several GB of activations and trial N+1's capture ``cudaMalloc``s them again — 0.2-1 s
    per trial, as much as 100 training steps of ResNet-50 (profiles/resnet_r2/trial_startup.md). Sharing the pool
    hands the blocks from one trial's graph straight to the next one's. ``MAGGY_B200_GRAPH_POOL=private`` restores
    the per-graph pools."""
    import os

    if os.environ.get("MAGGY_B200_GRAPH_POOL", "shared") == "private":
        return None
    key = torch.device(device).index or 0
    entry = _GRAPH_POOLS.get(key)
    if entry is None:
        pool = torch.cuda.graph_pool_handle()
        keep = torch.cuda.CUDAGraph()
        buf = torch.zeros(8, device=device)
        with torch.cuda.graph(keep, pool=pool):
            buf.add_(1.0)
        entry = _GRAPH_POOLS[key] = (pool, keep, buf)
    return entry[0]


class TrainStep:
    """``step(idx_host, tgt_host) -> loss`` (device scalar). Inputs may be pinned host or device tensors."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdamW, batch_shape, device,
                 use_cuda_graph: bool = True, warmup_iters: int = 3, input_dtype=torch.long,
                 loss_fn: Optional[Callable] = None, target_shape=None, memory_format=None):
        self.model, self.opt = model, optimizer
        self.device = torch.device(device)
        self.idx = torch.zeros(batch_shape, dtype=input_dtype, device=self.device)
        if memory_format is not None:
            self.idx = self.idx.contiguous(memory_format=memory_format)
        self.tgt = torch.zeros(batch_shape if target_shape is None else target_shape, dtype=torch.long,
                               device=self.device)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.loss_fn = loss_fn
        self.graph = None
        self.use_cuda_graph = use_cuda_graph
        self.warmup_iters = warmup_iters
        self.kernels_per_step = 0
        self._captured = False

    def _eager(self):
        self.opt.zero_grad()
        if self.loss_fn is not None:
            loss = self.loss_fn(self.model, self.idx, self.tgt)
        else:
            loss = self.model(self.idx, self.tgt)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def _capture(self):
        from ..ops import functional as F

        self.opt.graph_safe = True
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(self.warmup_iters):
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        before = F.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=_shared_graph_pool(self.device)):
            self._eager()
        self.kernels_per_step = F.launch_count() - before
        self._captured = True

    def load_batch(self, idx, tgt):
        self.idx.copy_(idx, non_blocking=True)
        self.tgt.copy_(tgt, non_blocking=True)

    def __call__(self, idx=None, tgt=None) -> torch.Tensor:
        if idx is not None:
            self.load_batch(idx, tgt)
        if self.use_cuda_graph:
            if not self._captured:
                with tracing.range("TrainStep.capture"):
                    self._capture()
            self.graph.replay()
        else:
            from ..ops import functional as F

            before = F.launch_count()
            self._eager()
            self.kernels_per_step = F.launch_count() - before
        return self.loss


def build_gpt2_trainer(cfg, batch_size: int, seq_len: int, device, world_size: int = 1, rank: int = 0,
                       group=None, lr: float = 3e-4, weight_decay: float = 0.1, use_cuda_graph: bool = True,
                       use_multicast="auto", seed: int = 1234):
    """Public convenience constructor used by bench.py / smoke(): GPT-2 + arena + fused AdamW + graph."""
    from ..models.gpt2 import GPT2

    torch.manual_seed(seed)
    model = GPT2(cfg)
    flat = FlatParams(model, device=device, world_size=world_size, rank=rank, group=group)
    opt = FusedAdamW(flat, lr=lr, weight_decay=weight_decay, use_multicast=use_multicast)
    step = TrainStep(model, opt, (batch_size, seq_len), device, use_cuda_graph=use_cuda_graph)
    return model, flat, opt, step


def build_trainer(model: torch.nn.Module, batch_shape, device, *, target_shape=None, input_dtype=torch.float32,
                  memory_format=None, loss_fn: Optional[Callable] = None, lr: float = 1e-3, betas=(0.9, 0.999),
                  weight_decay: float = 0.0, use_cuda_graph: bool = True, warmup_iters: int = 3,
                  world_size: int = 1, rank: int = 0, group=None):
    """Any ``nn.Module`` -> flat bf16 arena + fused AdamW + captured step (used by the HPO trial functions).

    Parameters whose gradients are produced by ATen autograd instead of the maggy_b200 wgrad kernels (BatchNorm,
    ``nn.Linear`` ...) are adopted into the arena on the first step (:meth:`FlatParams.adopt_autograd_grads`)."""
    model.to(device)
    flat = FlatParams(model, device=device, world_size=world_size, rank=rank, group=group)
    opt = FusedAdamW(flat, lr=lr, betas=betas, weight_decay=weight_decay)
    step = TrainStep(model, opt, batch_shape, device, use_cuda_graph=use_cuda_graph, warmup_iters=warmup_iters,
                     input_dtype=input_dtype, loss_fn=loss_fn, target_shape=target_shape, memory_format=memory_format)
    return flat, opt, step
