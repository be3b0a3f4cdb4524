"""Distribution-transparent data loading (reference: maggy/core/patching/dataloader.py:33-163).

``MaggyDataLoader`` is a drop-in ``torch.utils.data.DataLoader`` that shards the dataset over the ranks
with a ``DistributedSampler`` and delivers every batch on the GPU. Unlike the reference (synchronous
``tensor.cuda()`` from pageable memory per batch, dataloader.py:162) batches are staged in pinned memory
and copied on a side stream one batch ahead, so the H2D copy overlaps the previous step's compute.
``MaggyPetastormDataLoader`` reads a parquet dataset sharded by ``RANK`` / ``WORLD_SIZE`` with pyarrow
(petastorm itself is not required).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, DistributedSampler


def _to_device(obj, device, non_blocking=True):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, dict):
        return {k: _to_device(v, device, non_blocking) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device, non_blocking) for v in obj)
    return obj


def _to_cuda(data):
    """Recursively move a batch to the current CUDA device (reference: patching/dataloader.py:147-163)."""
    return _to_device(data, torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")


class MaggyDataLoader(DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, sampler=None, **kwargs):
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if sampler is None and dist.is_available() and dist.is_initialized() and hasattr(dataset, "__len__"):
            sampler = DistributedSampler(dataset, shuffle=shuffle)
            shuffle = False
        kwargs.setdefault("pin_memory", self._device is not None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle if sampler is None else False,
                         sampler=sampler, **kwargs)
        self._copy_stream = torch.cuda.Stream() if self._device is not None else None

    def __iter__(self):
        it = super().__iter__()
        if self._device is None:
            yield from it
            return
        nxt = None

        def stage(batch):
            with torch.cuda.stream(self._copy_stream):
                return _to_device(batch, self._device)

        for batch in it:
            staged = stage(batch)
            if nxt is not None:
                torch.cuda.current_stream().wait_stream(self._copy_stream)
                yield nxt
            nxt = staged
        if nxt is not None:
            torch.cuda.current_stream().wait_stream(self._copy_stream)
            yield nxt


class MaggyPetastormDataLoader:
    """Iterate a parquet dataset (path or list of files): this rank reads every ``WORLD_SIZE``-th row group."""

    def __init__(self, dataset, batch_size=1, columns=None, **_):
        import pyarrow.dataset as pads

        self.ds = pads.dataset(dataset, format="parquet") if isinstance(dataset, (str, list)) else dataset
        self.batch_size = batch_size
        self.columns = columns
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None

    def __iter__(self):
        frags = [f for i, f in enumerate(self.ds.get_fragments()) if i % self.world == self.rank]
        for frag in frags:
            for rb in frag.to_batches(batch_size=self.batch_size, columns=self.columns):
                batch = {name: torch.as_tensor(col.to_numpy(zero_copy_only=False)) for name, col in
                         zip(rb.schema.names, rb.columns)}
                yield _to_device(batch, self._device) if self._device is not None else batch

    def __len__(self):
        raise TypeError("a sharded parquet stream has no length")
