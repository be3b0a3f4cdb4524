"""`bench.py --impl reference`: the headline metric measured through the UNMODIFIED reference.

What runs: `maggy.experiment.lagom(train_fn, BaseConfig(...))` from the reference package installed in
`baseline/_ref` (pip --no-deps). The reference's python driver executes the oblivious `train_fn`
below in-process — exactly its stock single-run code path. `train_fn` is what a reference user writes for distributed
GPT-2 training: a plain PyTorch model (cuBLAS / ATen / SDPA kernels), wrapped with the reference's own
`get_maggy_ddp_wrapper` and optimised with its `MaggyZeroAdamW` when more than one rank runs (those are the classes
the reference's Spark torch executor hands to `train_fn`; under `torchrun` each rank applies them itself because the
Spark layer does not exist in this image). Nothing from `maggy_b200` is imported on this path.

The image has no TensorFlow / PySpark / skopt / statsmodels / DeepSpeed / fairscale, which the reference imports at
module scope; `baseline/refenv` supplies import-time stand-ins (documented there). The JSON line says so in
`reference_env`.
"""
from __future__ import annotations

import json
import math
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stock_gpt2(vocab, block, n_layer, n_head, n_embd):
    """GPT-2 in plain PyTorch (what the reference's users bring): nn.Linear / LayerNorm / SDPA / cross_entropy."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as tf

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.ln1, self.ln2 = nn.LayerNorm(n_embd), nn.LayerNorm(n_embd)
            self.qkv, self.proj = nn.Linear(n_embd, 3 * n_embd), nn.Linear(n_embd, n_embd)
            self.fc, self.fc2 = nn.Linear(n_embd, 4 * n_embd), nn.Linear(4 * n_embd, n_embd)

        def forward(self, x):
            B, T, C = x.shape
            q, k, v = self.qkv(self.ln1(x)).view(B, T, 3, n_head, C // n_head).unbind(2)
            y = tf.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
            x = x + self.proj(y.transpose(1, 2).reshape(B, T, C))
            return x + self.fc2(tf.gelu(self.fc(self.ln2(x)), approximate="tanh"))

    class GPT2(nn.Module):
        def __init__(self):
            super().__init__()
            self.vocab = (vocab + 127) // 128 * 128
            self.wte, self.wpe = nn.Embedding(self.vocab, n_embd), nn.Embedding(block, n_embd)
            self.blocks = nn.ModuleList([Block() for _ in range(n_layer)])
            self.ln_f = nn.LayerNorm(n_embd)
            self.apply(self._init)

        @staticmethod
        def _init(m):
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, std=0.02)
                if getattr(m, "bias", None) is not None:
                    nn.init.zeros_(m.bias)

        def forward(self, idx, targets):
            T = idx.shape[1]
            x = self.wte(idx) + self.wpe(torch.arange(T, device=idx.device))
            for b in self.blocks:
                x = b(x)
            logits = tf.linear(self.ln_f(x), self.wte.weight)  # tied head
            return tf.cross_entropy(logits.view(-1, logits.shape[-1]).float(), targets.view(-1))

    return GPT2


MODELS = {"gpt2-small": dict(vocab=50257, block=1024, n_layer=12, n_head=12, n_embd=768),
          "gpt2-tiny": dict(vocab=1000, block=128, n_layer=2, n_head=4, n_embd=128)}


def run(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def unavailable(why):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": why}))
        return 0

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "maggy")):
        return unavailable("baseline/_ref/maggy not installed (pip --no-index install failed or was not run)")
    try:
        import torch
    except Exception as e:  # noqa: BLE001
        return unavailable(f"torch import failed: {e}")
    if not torch.cuda.is_available():
        return unavailable("no CUDA device: the reference arm trains GPT-2 on the GPU")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    sys.path.insert(0, ref_dir)
    try:
        import refenv

        refenv.install()
        sys.modules.pop("maggy", None)  # never the repo's own `maggy` alias package
        import maggy as _ref_pkg

        if not os.path.abspath(_ref_pkg.__file__).startswith(os.path.abspath(ref_dir)):
            return unavailable(f"`import maggy` resolved to {_ref_pkg.__file__}, not to baseline/_ref")
        from maggy.config import BaseConfig
        from maggy.experiment import experiment as ref_experiment
    except Exception as e:  # noqa: BLE001
        return unavailable(f"reference import fails even with the stand-ins: {type(e).__name__}: {str(e)[:160]}")

    sys.path.insert(0, ROOT)
    from bench import ClockSampler  # engine-free helper (nvidia-smi sampling)

    spec = MODELS[args.model]
    B, T, K, W = args.batch, args.seq, args.steps, max(args.warmup, 3)
    holder = {}

    def train_fn():
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        notes = []
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=device)  # reference: torch_dist_executor.py:118-121
        torch.manual_seed(1234)
        Model = _stock_gpt2(**spec)
        if world > 1:
            try:
                from maggy.core.patching import get_maggy_ddp_wrapper

                model = get_maggy_ddp_wrapper(Model)()  # reference: patching/modules.py:44-63
                notes.append("module: reference MaggyDDPModuleWrapper")
            except Exception as e:  # noqa: BLE001
                from torch.nn.parallel import DistributedDataParallel as DDP

                model = DDP(Model().to(device), device_ids=[local_rank])
                notes.append(f"module: torch DDP (reference wrapper failed on torch {torch.__version__}: {str(e)[:80]})")
            try:
                from maggy.core.patching import MaggyZeroAdamW

                opt = MaggyZeroAdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1)
                notes.append("optimizer: reference MaggyZeroAdamW")
            except Exception as e:  # noqa: BLE001
                from torch.distributed.optim import ZeroRedundancyOptimizer

                opt = ZeroRedundancyOptimizer(model.parameters(), optimizer_class=torch.optim.AdamW, lr=3e-4,
                                              betas=(0.9, 0.95), weight_decay=0.1)
                notes.append("optimizer: torch ZeroRedundancyOptimizer(AdamW) — the class MaggyZeroAdamW subclasses "
                             f"(its ctor failed on torch {torch.__version__}: {str(e)[:80]})")
        else:
            model = Model().to(device)
            opt = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1)
            notes.append("single rank: torch.optim.AdamW on the user's module (reference wraps nothing at world=1)")

        g = torch.Generator().manual_seed(1000 + rank)
        pool = [(torch.randint(0, spec["vocab"], (B, T), generator=g).pin_memory(),
                 torch.randint(0, spec["vocab"], (B, T), generator=g).pin_memory()) for _ in range(4)]
        idx = torch.zeros((B, T), dtype=torch.long, device=device)
        tgt = torch.zeros((B, T), dtype=torch.long, device=device)

        def step(batch=None):
            if batch is not None:
                idx.copy_(batch[0], non_blocking=True)
                tgt.copy_(batch[1], non_blocking=True)
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = model(idx, tgt)
            loss.backward()
            opt.step()
            return loss

        def sync_all():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(device)

        step(pool[0])
        for _ in range(W):
            step()
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(K):
            loss = step()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        final_loss = float(loss.item())
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        f0.record()
        for i in range(K):
            float(step(pool[i % len(pool)]).item())  # pinned H2D of the batch + D2H of the loss every step
        f1.record()
        sync_all()
        ms_e2e = f0.elapsed_time(f1)
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        holder.update(ms=float(t[0]), ms_e2e=float(t[1]), clocks=clocks, final_loss=final_loss, notes=notes,
                      h2d=sum(x.numel() * x.element_size() for x in pool[0]),
                      params=sum(p.numel() for p in model.parameters()))
        return {"tokens_per_sec": B * T * world * K / (float(t[0]) / 1e3)}

    # the reference writes its experiment log under the working directory: one scratch directory per rank
    os.chdir(tempfile.mkdtemp(prefix=f"maggy_ref_rank{rank}_"))
    try:
        import contextlib

        with contextlib.redirect_stdout(sys.stderr):  # the reference prints its own progress lines: keep stdout = JSON
            ref_experiment.lagom(train_fn, BaseConfig(name="bench_reference"))
    except Exception as e:  # noqa: BLE001
        return unavailable(f"reference lagom() raised {type(e).__name__}: {str(e)[:200]}")
    if not holder or not math.isfinite(holder.get("ms", float("nan"))):
        return unavailable("reference lagom() returned without running train_fn")
    tokens = B * T * world * K
    out = {
        "metric": "dist-train tokens/sec (GPT-2 small bf16 data-parallel, fused reduce-scatter+Adam)",
        "value": tokens / (holder["ms"] / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": holder["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic random tokens, random-init weights", "impl": "reference",
        "config": {"model": args.model, "params": holder["params"], "global_batch": B * world, "seq_len": T,
                   "per_gpu_batch": B, "parallelism": f"dp{world}" + ("+ddp+zero1" if world > 1 else ""),
                   "attention": "sdpa", "cuda_graph": False,
                   "l2": "per-step working set (activations + logits, several GB) >> 126 MB L2; no explicit flush"},
        "clocks": holder["clocks"],
        "e2e": {"value": tokens / (holder["ms_e2e"] / 1e3), "unit": "tokens/s", "ms_per_step": holder["ms_e2e"] / K,
                "h2d_bytes_per_step": holder["h2d"], "d2h_bytes_per_step": 4},
        "gpu_launches": 0,
        "final_loss": holder["final_loss"],
        "reference_env": {"package": "baseline/_ref (maggy 1.1.2, unmodified)", "entry": "maggy.experiment.lagom + BaseConfig",
                          "driver": "reference python (non-Spark) driver, train_fn in-process on every torchrun rank",
                          "stand_ins": "tensorflow (tensorboard stub), placeholders for skopt/statsmodels/deepspeed/"
                                       "fairscale/petastorm/hops (baseline/refenv)",
                          "wiring": holder["notes"]},
    }
    if rank == 0:
        print(json.dumps(out))
    try:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass
    return 0
