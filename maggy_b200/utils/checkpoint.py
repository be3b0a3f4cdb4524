"""Checkpoint / resume of the flat-arena training state (additive capability, SURVEY.md §5.4 — the reference has none:
promoted ASHA / Hyperband trials retrain from scratch).

One directory per checkpoint::

    <dir>/meta.json              step, world size, arena geometry, user extras
    <dir>/params.pt              the full bf16 parameter arena (written by rank 0)
    <dir>/optim_rank<r>.pt       rank r's fp32 master / m / v shard

The optimizer shards are the source of truth (ZeRO layout: every rank saves and loads only its own slice); the bf16
arena is stored once so that a restart does not need a collective to rebuild it.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch


def _geometry(flat) -> dict:
    return {"numel": int(flat.numel), "shard_numel": int(flat.shard_numel), "world_size": int(flat.world_size),
            "decay_boundary": int(flat.decay_boundary),
            "entries": [[n, int(o), int(k)] for n, _, o, k in flat.entries]}


def save_training_state(directory: str, flat, opt, extra: Optional[dict] = None) -> str:
    """Write this rank's part of the checkpoint (call on every rank; rank 0 also writes the shared files)."""
    os.makedirs(directory, exist_ok=True)
    sd = opt.state_dict()
    torch.save({"step": sd["step"], "m": sd["m"].cpu(), "v": sd["v"].cpu(), "master": sd["master"].cpu()},
               os.path.join(directory, f"optim_rank{flat.rank}.pt"))
    if flat.rank == 0:
        torch.save(flat.params.detach().cpu(), os.path.join(directory, "params.pt"))
    if flat.world_size > 1 and torch.distributed.is_initialized():
        # meta.json = "checkpoint complete": every rank's optimizer shard must be on disk before rank 0 writes it
        torch.distributed.barrier(group=getattr(flat, "group", None))
    if flat.rank == 0:
        meta = {"format": 1, "step": int(sd["step"]), "geometry": _geometry(flat), "extra": extra or {}}
        tmp = os.path.join(directory, "meta.json.tmp")
        with open(tmp, "w") as f:
            json.dump(meta, f)
        os.replace(tmp, os.path.join(directory, "meta.json"))  # meta.json appears last: its presence = complete
    return directory


def load_training_state(directory: str, flat, opt) -> dict:
    """Restore arena + optimizer state saved by :func:`save_training_state`; returns the ``extra`` dict."""
    meta = json.load(open(os.path.join(directory, "meta.json")))
    geo = meta["geometry"]
    mine = _geometry(flat)
    for key in ("numel", "shard_numel", "world_size", "decay_boundary"):
        if geo[key] != mine[key]:
            raise ValueError(f"checkpoint {directory} was written for {key}={geo[key]}, this run has {mine[key]}")
    if [e[0] for e in geo["entries"]] != [e[0] for e in mine["entries"]]:
        raise ValueError(f"checkpoint {directory}: parameter names / order differ from this model")
    shard = torch.load(os.path.join(directory, f"optim_rank{flat.rank}.pt"), map_location="cpu")
    opt.load_state_dict({"step": shard["step"], "m": shard["m"].to(opt.m.device), "v": shard["v"].to(opt.v.device),
                         "master": shard["master"].to(flat.master.device)})
    if getattr(opt, "step_tensor", None) is not None:
        opt.step_tensor.fill_(int(shard["step"]))  # the graph-captured step counter lives on the device
    with torch.no_grad():
        flat.params.copy_(torch.load(os.path.join(directory, "params.pt"), map_location="cpu").to(flat.params.device))
    return meta.get("extra", {})


def latest(directory: str) -> Optional[str]:
    """Newest complete checkpoint under ``directory`` (sub-directories named by step), or None."""
    if not os.path.isdir(directory):
        return None
    done = [d for d in os.listdir(directory) if os.path.exists(os.path.join(directory, d, "meta.json"))]
    if not done:
        return None
    return os.path.join(directory, max(done, key=lambda d: json.load(open(os.path.join(directory, d, "meta.json")))["step"]))
