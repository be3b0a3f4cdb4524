"""K7 micro-benchmark: maggy_b200 conv2d (implicit GEMM / GEMM / im2col routes) vs cuDNN (torch, channels_last bf16) on
the ResNet-50 layer shapes, forward and forward+backward, plus one full ResNet-50 training step for both
implementations. CUDA-event timed, L2 flushed between iterations.
Usage: python scripts/bench_conv.py [--batch 64] [--out profiles/conv_bench.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.models.resnet import resnet50  # noqa: E402
from maggy_b200.ops.conv import Conv2d  # noqa: E402


def timeit(fn, flush, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = "cuda:0"
    B = args.batch
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    layers = [  # name, H, C, K, k, stride, pad
        ("stem 7x7/2 3->64 @224", 224, 3, 64, 7, 2, 3),
        ("l1 1x1 64->64 @56", 56, 64, 64, 1, 1, 0),
        ("l1 3x3 64->64 @56", 56, 64, 64, 3, 1, 1),
        ("l1 1x1 64->256 @56", 56, 64, 256, 1, 1, 0),
        ("l2 3x3/2 128->128 @56", 56, 128, 128, 3, 2, 1),
        ("l2 3x3 128->128 @28", 28, 128, 128, 3, 1, 1),
        ("l2 1x1 128->512 @28", 28, 128, 512, 1, 1, 0),
        ("l3 3x3 256->256 @14", 14, 256, 256, 3, 1, 1),
        ("l3 1x1 1024->256 @14", 14, 1024, 256, 1, 1, 0),
        ("l4 3x3 512->512 @7", 7, 512, 512, 3, 1, 1),
        ("l4 1x1/2 1024->2048 @14", 14, 1024, 2048, 1, 2, 0),
    ]
    rows = []
    for name, H, C, K, k, stride, pad in layers:
        ours = Conv2d(C, K, k, stride=stride, padding=pad).to(dev)
        ref = torch.nn.Conv2d(C, K, k, stride=stride, padding=pad, bias=False).to(dev).to(torch.bfloat16)
        ref = ref.to(memory_format=torch.channels_last)
        x = torch.randn(B, C, H, H, device=dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        need_dx = C != 3
        xo = x.clone().requires_grad_(need_dx)
        xr = x.clone().requires_grad_(need_dx)
        yo = ours(xo)
        g = torch.randn_like(yo)

        def fb(mod, inp):
            mod.weight.grad = None
            if inp.grad is not None:
                inp.grad = None
            mod(inp).backward(g)

        with torch.no_grad():
            t_of = timeit(lambda: ours(x), flush)
            t_rf = timeit(lambda: ref(x), flush)
        t_ob = timeit(lambda: fb(ours, xo), flush)
        t_rb = timeit(lambda: fb(ref, xr), flush)
        P = (H + 2 * pad - k) // stride + 1
        fl = 2.0 * B * P * P * K * C * k * k
        row = {"layer": name, "batch": B, "ours_fwd_ms": t_of, "cudnn_fwd_ms": t_rf, "ours_fwdbwd_ms": t_ob,
               "cudnn_fwdbwd_ms": t_rb, "ours_fwd_tflops": fl / t_of / 1e9, "cudnn_fwd_tflops": fl / t_rf / 1e9}
        rows.append(row)
        print(f"{name:28s} fwd ours {t_of:7.3f} ms ({row['ours_fwd_tflops']:6.0f} TF) cudnn {t_rf:7.3f} ms "
              f"({row['cudnn_fwd_tflops']:6.0f} TF) | fwd+bwd ours {t_ob:7.3f} cudnn {t_rb:7.3f} "
              f"| speedup fwd {t_rf / t_of:.2f} f+b {t_rb / t_ob:.2f}", flush=True)
    # full model step
    res = {}
    for impl in ("native", "torch"):
        torch.manual_seed(0)
        model = resnet50(1000, conv_impl=impl).to(dev).to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
        x = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (B,), device=dev)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(model(x), y)
            loss.backward()
            opt.step()

        res[impl] = timeit(step, flush, iters=10, warm=4)
        print(f"ResNet-50 train step (batch {B}) conv_impl={impl}: {res[impl]:.2f} ms  "
              f"({B / res[impl] * 1e3:.0f} img/s)", flush=True)
        del model, opt
    rows.append({"resnet50_step_ms": res, "batch": B})
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
