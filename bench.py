#!/usr/bin/env python
"""Headline benchmark: GPT-2 small bf16 data-parallel training throughput (tokens/s) on N B200s.

Contract (see task statement): ``python bench.py --gpus N --steps K --warmup W`` — for N > 1 it is
launched under ``torch.distributed.run`` with one rank per GPU. Prints ONE JSON line on rank 0.

Both arms go through the product's public entry point: every torchrun rank calls ``experiment.lagom(train_fn,
BaseConfig(...))`` (ours: ``maggy_b200.experiment.lagom``; ``--impl reference``: the unmodified reference's
``maggy.experiment.lagom`` from ``baseline/_ref``, see ``baseline/reference_arm.py`` and DESIGN.md section 5) and the
oblivious ``train_fn`` builds the model, trains and reports through ``reporter.broadcast``.

* ``value``   device-timed (CUDA events, max over ranks) whole-job tokens/s of the full training step:
              forward + backward + fused reduce-scatter/AdamW/all-gather, replayed as a CUDA graph.
* ``e2e``     the same metric through the public API including, every step, the pinned-host -> device copy of that
              step's batch, a device -> host read of the loss and ``reporter.broadcast(loss, step)``.
* guards      at N > 1, before timing: parameters bit-identical on every rank after the warm-up steps (the fused
              kernel's all-gather epilogue) and a finite loss; the line carries ``grad_dtype``.
* ``--config asha``  BASELINE config 2 (ASHA over ResNet-50, one trial per GPU): trials/hour through
              ``lagom(HyperparameterOptConfig)`` — ``bench/hpo_asha_resnet.py``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch"])
    ap.add_argument("--model", default="gpt2-small")
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU per step")
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--attn", default="", choices=["", "sdpa", "native"],
                    help="attention implementation of the model (default: MAGGY_B200_ATTN or native = K10)")
    ap.add_argument("--no-multicast", action="store_true", help="force the peer load/store variant of the fused kernel")
    ap.add_argument("--multicast", action="store_true", help="force the NVSwitch multimem variant (default: auto)")
    ap.add_argument("--config", default="gpt2-dp", choices=["gpt2-dp", "asha"],
                    help="gpt2-dp: BASELINE config 3 (headline); asha: BASELINE config 2, trials/hour")
    ap.add_argument("--trials", type=int, default=64, help="--config asha: number of trials")
    ap.add_argument("--steps-per-budget", type=int, default=200, help="--config asha: training steps per budget unit")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def _pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())

        self.thread = threading.Thread(target=_pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args):
    """Run the UNMODIFIED reference through its own public API (baseline/reference_arm.py), or say why it cannot."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import reference_arm as ra
    except Exception as e:  # noqa: BLE001
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"baseline/reference_arm.py failed to import: {e}"}))
        return 0
    return ra.run(args)


def asha_config(args):
    """BASELINE config 2 through ``lagom(HyperparameterOptConfig(optimizer=Asha))``: one line of JSON, trials/hour."""
    cmd = [sys.executable, os.path.join(ROOT, "bench", "hpo_asha_resnet.py"), "--workers", str(args.gpus),
           "--trials", str(args.trials), "--steps-per-budget", str(args.steps_per_budget), "--batch", str(min(args.batch, 64))]
    if args.impl == "reference":
        # the reference's HPO driver needs Spark (maggy/config/hyperparameter_optimization.py:78-81): labelled proxy
        cmd += ["--proxy-reference-rpc", "--engine", "eager"]
    return subprocess.call(cmd)


def main():
    args = parse_args()
    if args.config == "asha":
        return asha_config(args)
    if args.impl == "reference":
        return reference_arm(args)
    if args.impl == "ours":
        return ours_through_lagom(args)
    return run_training(args, reporter=None)


def ours_through_lagom(args):
    """The product path: ``experiment.lagom(train_fn, BaseConfig)`` on every torchrun rank (the same shape as the
    reference arm: its python driver also runs the oblivious train_fn in-process on each rank)."""
    import tempfile

    rank = int(os.environ.get("RANK", "0"))
    # every torchrun rank runs its own in-process experiment: each needs its own artefact directory
    base = os.environ.get("MAGGY_LOG_DIR")
    if base:
        os.makedirs(base, exist_ok=True)
    os.environ["MAGGY_LOG_DIR"] = tempfile.mkdtemp(prefix=f"maggy_b200_bench_rank{rank}_", dir=base or None)
    from maggy_b200 import experiment
    from maggy_b200.config import BaseConfig

    def train_fn(reporter):
        return run_training(args, reporter=reporter)

    import contextlib
    import io

    sink = io.StringIO()  # the driver prints its own summary; the contract is ONE JSON line on stdout
    with contextlib.redirect_stdout(sink if rank == 0 else io.StringIO()):
        result = experiment.lagom(train_fn, BaseConfig(name="bench_gpt2_dp", description="BASELINE config 3"))
    line = getattr(run_training, "json_line", None)
    if rank == 0:
        if line is None:
            print(json.dumps({"impl": "ours", "error": "train_fn produced no result", "lagom_result": str(result)[:200]}))
            return 1
        print(line)
    return 0


def run_training(args, reporter=None):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from maggy_b200.models.gpt2 import GPT2Config
    from maggy_b200.ops import functional as F

    cfg = GPT2Config.named(args.model)
    if args.attn:
        cfg.attn_impl = args.attn
    B, T = args.batch, args.seq
    K, W = args.steps, max(args.warmup, 3)

    if args.impl == "ours":
        from maggy_b200.parallel.trainer import build_gpt2_trainer

        model, flat, opt, step = build_gpt2_trainer(
            cfg, B, T, device, world_size=world, rank=rank, use_cuda_graph=not args.no_graph,
            use_multicast=False if args.no_multicast else (True if args.multicast else "auto"))
        uses_mc = opt.uses_multicast
    else:
        step, uses_mc = build_torch_baseline(cfg, B, T, device, world, rank)

    # a pool of pinned host batches (synthetic tokens): every step copies a different one
    g = torch.Generator().manual_seed(1000 + rank)
    pool = [(torch.randint(0, cfg.vocab_size, (B, T), generator=g).pin_memory(),
             torch.randint(0, cfg.vocab_size, (B, T), generator=g).pin_memory()) for _ in range(4)]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---------------- warm-up (includes CUDA-graph capture) ----------------
    step.load_batch(*pool[0])
    for i in range(W):
        step()
    sync_all()

    # ---------------- correctness guards (untimed) --------------------------
    guards = {"grad_dtype": str(flat.grads.dtype).replace("torch.", "") if args.impl == "ours" else "float32"}
    if args.impl == "ours":
        guards["loss_finite_after_warmup"] = bool(torch.isfinite(step.loss).item())
        if world > 1:
            # the fused kernel's all-gather epilogue must leave every rank with the same bf16 parameters, bit for bit
            words = flat.params.detach().view(torch.int16).to(torch.int64)
            chk = torch.stack([words.sum(), (words * (torch.arange(words.numel(), device=device) % 8191 + 1)).sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            guards["params_bit_identical_across_ranks"] = bool((lo == hi).all().item())
            if not guards["params_bit_identical_across_ranks"]:
                raise RuntimeError("bench guard: parameters differ between ranks after the warm-up steps")
        if not guards["loss_finite_after_warmup"]:
            raise RuntimeError("bench guard: non-finite loss after warm-up")

    # ---------------- device-timed region: exactly K steps -----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches_before = F.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(K):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(step.loss.item())

    # ---------------- end-to-end region through the public API -------------
    h2d = sum(t.numel() * t.element_size() for t in pool[0])
    d2h = 4
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    f0.record()
    host_losses = []
    for i in range(K):
        idx, tgt = pool[i % len(pool)]
        loss = step(idx, tgt)          # pinned H2D copy + graph replay
        host_losses.append(float(loss.item()))  # D2H read of the step's result
        if reporter is not None:
            reporter.broadcast(host_losses[-1], i)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    wall_e2e_ms = (time.perf_counter() - t_wall0) * 1e3

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    tokens = B * T * world * K
    value = tokens / (ms / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)
    if args.impl == "ours":
        launches = step.kernels_per_step * K if not args.no_graph else F.launch_count() - launches_before
    else:
        launches = 0
    out = {
        "metric": "dist-train tokens/sec (GPT-2 small bf16 data-parallel, fused reduce-scatter+Adam)",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic random tokens, random-init weights",
        "impl": args.impl,
        "config": {"attention": (args.attn or os.environ.get("MAGGY_B200_ATTN", "native")) if args.impl == "ours" else "sdpa",
                   "model": args.model, "params": sum(p.numel() for p in model.parameters()) if args.impl == "ours" else None,
                   "global_batch": B * world, "seq_len": T, "per_gpu_batch": B,
                   "parallelism": f"dp{world}" + ("+zero1-fused-rs-adam" if args.impl == "ours" else "+ddp+zero1"),
                   "nvls_multicast": bool(uses_mc), "cuda_graph": not args.no_graph,
                   "l2": "per-step working set (activations + logits, several GB) >> 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "final_loss": final_loss,
        "guards": guards,
        "api": "maggy_b200.experiment.lagom(train_fn, BaseConfig) per rank; reporter.broadcast(loss, step) in the e2e "
               "region" if reporter is not None else "engine-internal proxy (no lagom)",
        "mfu_vs_measured_bf16_peak": None,
    }
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if args.impl == "ours":
            fl = model.flops_per_token(T) * B * T / (ms / K / 1e3)  # per GPU
            out["mfu_vs_measured_bf16_peak"] = fl / (peaks["bf16_tflops_sustained"] * 1e12)
            out["model_tflops_per_gpu"] = fl / 1e12
    except Exception:
        pass
    run_training.json_line = json.dumps(out)
    if rank == 0 and reporter is None:
        print(run_training.json_line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return {"tokens_per_sec": value} if reporter is not None else 0


def build_torch_baseline(cfg, B, T, device, world, rank):
    """Stock-library arm (proxy of the reference's mechanisms): DDP + ZeroRedundancyOptimizer(AdamW)
    over NCCL, cuBLAS/ATen kernels, bf16 autocast — what maggy's torch patching wires up
    (maggy/core/patching/modules.py:44-63, optim.py:28-45)."""
    import torch
    import torch.distributed as dist

    from maggy_b200.models.gpt2 import TorchGPT2

    torch.manual_seed(1234)
    model = TorchGPT2(cfg).to(device)
    if world > 1:
        from torch.distributed.optim import ZeroRedundancyOptimizer
        from torch.nn.parallel import DistributedDataParallel as DDP

        ddp = DDP(model, device_ids=[device.index])
        opt = ZeroRedundancyOptimizer(ddp.parameters(), optimizer_class=torch.optim.AdamW, lr=3e-4,
                                      betas=(0.9, 0.95), weight_decay=0.1, fused=True)
    else:
        ddp = model
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1, fused=True)

    class Step:
        def __init__(self):
            self.idx = torch.zeros((B, T), dtype=torch.long, device=device)
            self.tgt = torch.zeros((B, T), dtype=torch.long, device=device)
            self.loss = torch.zeros((), device=device)
            self.kernels_per_step = 0

        def load_batch(self, idx, tgt):
            self.idx.copy_(idx, non_blocking=True)
            self.tgt.copy_(tgt, non_blocking=True)

        def __call__(self, idx=None, tgt=None):
            if idx is not None:
                self.load_batch(idx, tgt)
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = ddp(self.idx, self.tgt)
            loss.backward()
            opt.step()
            self.loss.copy_(loss.detach())
            return self.loss

    return Step(), False


if __name__ == "__main__":
    sys.exit(main())
