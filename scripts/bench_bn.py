"""Fused NHWC BatchNorm(+ReLU,+add) vs ATen on the ResNet-50 layer shapes (batch 64) and the full ResNet-50 step.
Per-layer numbers are CUDA-graph replays (kernel time, no host launch overhead) with L2 flushed between replays;
roofline = bytes that must move / measured copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_gemm import timeit  # noqa: E402
from maggy_b200.ops.batchnorm import BatchNorm2d  # noqa: E402

dev = "cuda:0"
PEAK = 6576e9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:  # noqa: BLE001
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B = int(os.environ.get("BN_B", 64))
rows = []


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


for name, C, HW, relu, res in [("stem bn+relu 64@112", 64, 112, True, False), ("l1 bn+relu 64@56", 64, 56, True, False),
                               ("l1 bn+add+relu 256@56", 256, 56, True, True), ("l2 bn+relu 128@28", 128, 28, True, False),
                               ("l2 bn+add+relu 512@28", 512, 28, True, True), ("l3 bn+add+relu 1024@14", 1024, 14, True, True),
                               ("l4 bn+add+relu 2048@7", 2048, 7, True, True), ("l4 bn 2048@7 (downsample)", 2048, 7, False, False)]:
    x = torch.randn(B, C, HW, HW, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn_like(x).requires_grad_(True) if res else None
    g = torch.randn(B, C, HW, HW, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ours = BatchNorm2d(C, relu=relu).to(dev)
    ref = torch.nn.BatchNorm2d(C).to(dev)

    def f_ours():
        y = ours(x, r)
        y.backward(g)
        x.grad = None
        if r is not None:
            r.grad = None

    def f_ref():
        y = ref(x)
        if res:
            y = y + r
        if relu:
            y = torch.relu(y)
        y.backward(g)
        x.grad = None
        if r is not None:
            r.grad = None

    t_o = timeit(graphed(f_ours), flush)
    t_r = timeit(graphed(f_ref), flush)
    nbytes = x.numel() * 2
    # fwd: read x twice (+res), write y; bwd: read dy,y,x twice, write dx (+dz)
    ideal = nbytes * ((3 + (1 if res else 0)) + (2 * (3 if relu else 2) + 1 + (1 if res and relu else 0)))
    roof = ideal / PEAK * 1e3
    rows.append({"layer": name, "ours_ms": t_o, "aten_ms": t_r, "roofline_ms": roof})
    print(f"{name:30s} fwd+bwd ours {t_o:7.3f} ms  aten {t_r:7.3f} ms  speedup {t_r / t_o:5.2f}x   roofline {roof:6.3f} ms "
          f"({roof / t_o * 100:4.1f} %)", flush=True)

if os.environ.get("BN_STEP", "1") != "0":
    from maggy_b200.models.resnet import resnet50

    for label, impl, bn_env in (("native conv + fused norm", "native", "native"), ("native conv + ATen norm", "native", "aten"),
                                ("cuDNN conv + ATen norm", "torch", "aten")):
        os.environ["MAGGY_B200_BN"] = bn_env
        torch.manual_seed(0)
        model = resnet50(1000, conv_impl=impl).to(dev).to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
        xx = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
        yy = torch.randint(0, 1000, (B,), device=dev)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(model(xx), yy)
            loss.backward()
            opt.step()

        t = timeit(step, flush, iters=10, warm=4)
        rows.append({"step": label, "ms": t, "img_per_s": B / t * 1e3})
        print(f"ResNet-50 train step (batch {B}, eager) {label:28s}: {t:7.2f} ms  ({B / t * 1e3:6.0f} img/s)", flush=True)
        del model, opt
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
