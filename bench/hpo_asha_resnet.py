#!/usr/bin/env python
"""BASELINE.json config 2: ASHA hyperparameter search over ResNet-50 on synthetic ImageNet-shaped data,
one trial per GPU, all GPUs busy. Reports trials/hour plus the control-plane numbers that bound it:
inter-trial dead time (FINAL -> next trial assigned) and METRIC -> STOP latency.

    python bench/hpo_asha_resnet.py --workers 8 --trials 32 --steps-per-budget 10 --batch 64
    python bench/hpo_asha_resnet.py --workers 2 --model tiny --cpu         # plumbing check without GPUs

``--proxy-reference-rpc`` makes the workers emulate the reference's control-plane constants (metrics
shipped once per 1 s heartbeat, 1 s GET poll after FINAL; maggy/core/rpc.py:716-748) — a labelled proxy of
the reference's mechanism, not the reference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_train_fn(model_name, batch, steps_per_budget, image, classes, engine="graph"):
    """engine="graph": native tcgen05 convolutions + flat-arena fused AdamW, the whole step replayed as one CUDA
    graph (the maggy_b200 path). engine="eager": cuDNN convolutions + torch SGD, op-by-op (what the reference's
    oblivious train_fn does on this hardware)."""

    def train_fn_graph(hparams, reporter):
        import torch

        from maggy_b200.models.resnet import resnet50
        from maggy_b200.parallel.trainer import build_trainer

        import time as _t

        t0 = _t.time()
        dev = torch.device("cuda", 0)
        torch.manual_seed(0)
        with torch.device(dev):  # parameters are initialised ON the GPU: no CPU-side randn / host-to-device copies
            model = resnet50(classes, dropout=hparams["dropout"])
        model = model.to(memory_format=torch.channels_last)
        torch.cuda.synchronize()
        t_model = _t.time() - t0

        def loss_fn(m, x, y):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return torch.nn.functional.cross_entropy(m(x), y)

        _, _, step_fn = build_trainer(model, (batch, 3, image, image), dev, target_shape=(batch,),
                                      memory_format=torch.channels_last, loss_fn=loss_fn, lr=hparams["lr"] * 0.01,
                                      betas=(hparams["momentum"], 0.999), weight_decay=hparams["wd"], warmup_iters=2)
        x = torch.randn(batch, 3, image, image, device=dev).to(memory_format=torch.channels_last)
        y = torch.randint(0, classes, (batch,), device=dev)
        step_fn.load_batch(x, y)
        steps = int(hparams.get("budget", 1)) * steps_per_budget
        loss = None
        t1 = _t.time()
        step_fn()  # first call: warm-up iterations + graph capture (they are training steps too)
        torch.cuda.synchronize()
        t_capture = _t.time() - t1
        t2 = _t.time()
        for step in range(1, steps):
            loss = step_fn()
            if step % 2 == 1 or step == steps - 1:
                reporter.broadcast(-loss.detach().float(), step)  # published from the GPU (K8), no .item()
        final = -float(step_fn.loss if loss is None else loss)
        reporter.log(f"TIMING model={t_model:.3f}s trainer+capture={t_capture + (t1 - t0 - t_model):.3f}s "
                     f"steps({steps - 1})={_t.time() - t2:.3f}s", False)
        return final

    def train_fn(hparams, reporter):
        import torch

        from maggy_b200.models.resnet import resnet18ish, resnet50

        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        if engine == "graph" and dev.type == "cuda" and model_name == "resnet50":
            return train_fn_graph(hparams, reporter)
        torch.manual_seed(0)
        impl = "torch" if engine == "eager" else "native"
        with torch.device(dev):  # initialise on the device (see train_fn_graph)
            model = (resnet50(classes, dropout=hparams["dropout"], conv_impl=impl) if model_name == "resnet50"
                     else resnet18ish(classes, conv_impl=impl))
        if dev.type == "cuda":
            model = model.to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(model.parameters(), lr=hparams["lr"], momentum=hparams["momentum"],
                              weight_decay=hparams["wd"], nesterov=True)
        x = torch.randn(batch, 3, image, image, device=dev)
        y = torch.randint(0, classes, (batch,), device=dev)
        if dev.type == "cuda":
            x = x.to(memory_format=torch.channels_last)
        steps = int(hparams.get("budget", 1)) * steps_per_budget
        loss = None
        for step in range(steps):
            opt.zero_grad(set_to_none=True)
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
            if step % 2 == 1 or step == steps - 1:
                # the loss stays on the device: the reporter publishes it from the GPU (K8), no .item()
                reporter.broadcast(-loss.detach().float(), step)
        return -float(loss)

    return train_fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=0, help="0 = one per visible GPU")
    ap.add_argument("--trials", type=int, default=32)
    ap.add_argument("--steps-per-budget", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "tiny"])
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--proxy-reference-rpc", action="store_true")
    ap.add_argument("--es", default="median", choices=["median", "none"])
    ap.add_argument("--engine", default="graph", choices=["graph", "native-eager", "eager"],
                    help="graph (default): K7 convs + fused AdamW, whole step as one CUDA graph; eager: cuDNN + torch "
                         "SGD, op by op (what a reference user's train_fn does); native-eager: K7 convs, op by op")
    args = ap.parse_args()
    if args.proxy_reference_rpc:
        os.environ["MAGGY_B200_PROXY_REFERENCE_RPC"] = "1"
    if args.cpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import torch

    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig
    from maggy_b200.optimizer import Asha

    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    workers = args.workers or max(1, ngpu)
    image = args.image if args.model == "resnet50" else 32
    classes = 1000 if args.model == "resnet50" else 10
    sp = Searchspace(lr=("DOUBLE", [0.01, 0.4]), momentum=("DOUBLE", [0.8, 0.99]), wd=("DOUBLE", [1e-5, 1e-3]),
                     dropout=("DOUBLE", [0.0, 0.5]))
    cfg = HyperparameterOptConfig(num_trials=args.trials, optimizer=Asha(reduction_factor=2, resource_min=1, resource_max=4),
                                  searchspace=sp, direction="max", es_policy=args.es, es_interval=1, es_min=4,
                                  name="asha_resnet50", num_workers=workers, seed=0)
    train_fn = make_train_fn(args.model, args.batch, args.steps_per_budget, image, classes, engine=args.engine)
    from maggy_b200.experiment import experiment as exp_mod

    t0 = time.time()
    driver_holder = {}
    orig = exp_mod.lagom_driver

    def spy(config, app_id, run_id):
        d = orig(config, app_id, run_id)
        driver_holder["d"] = d
        return d

    exp_mod.lagom_driver = spy
    try:
        res = experiment.lagom(train_fn, cfg)
    finally:
        exp_mod.lagom_driver = orig
    wall = time.time() - t0
    d = driver_holder["d"]
    gaps = sorted(d.inter_trial_gaps)
    lats = sorted(d.stop_latencies)
    med = lambda v: v[len(v) // 2] if v else None  # noqa: E731
    out = {
        "metric": "trials/hour (ASHA, ResNet-50 synthetic, 1 GPU/trial)", "value": res["num_trials"] / wall * 3600.0,
        "unit": "trials/hour", "n_gpus": ngpu, "workers": workers, "trials": res["num_trials"],
        "early_stopped": res["early_stopped"], "wall_s": wall, "experiment_s": d.duration / 1e3,
        "inter_trial_gap_ms_median": None if not gaps else med(gaps) * 1e3,
        "inter_trial_gap_ms_max": None if not gaps else gaps[-1] * 1e3,
        "report_to_stop_latency_ms_median": None if not lats else med(lats) * 1e3,
        "proxy_reference_rpc": bool(args.proxy_reference_rpc),
        "device_scheduler": (dict(d.device_scheduler_stats or {}, decisions_applied=d.device_decisions,
                                  decision_fallbacks=d.device_decision_fallbacks, handoffs_adopted=d.device_handoffs,
                                  final_to_decision_ms_median=(sorted(d.device_decision_latencies)[len(d.device_decision_latencies) // 2]
                                                               if d.device_decision_latencies else None))
                             if getattr(d, "device_scheduler_stats", None) else None),
        "config": {"model": args.model, "batch": args.batch, "image": image, "steps_per_budget": args.steps_per_budget,
                   "optimizer": "asha(rf=2, min=1, max=4)", "es_policy": args.es, "engine": args.engine},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
