"""K7 conv2d kernels vs a plain PyTorch fp32 reference (needs a B200)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_conv(x_nhwc, w_krsc, stride, padding):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_krsc.float().permute(0, 3, 1, 2)
    return torch.nn.functional.conv2d(x, w, None, stride, padding).permute(0, 2, 3, 1)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


IMPLICIT_SHAPES = [
    # N, H, W, C, Kout, R, S, pad
    (4, 14, 14, 64, 64, 3, 3, 1),
    (2, 56, 56, 64, 128, 3, 3, 1),
    (8, 7, 7, 128, 256, 3, 3, 1),
    (3, 28, 28, 128, 128, 3, 3, 1),
    (2, 9, 13, 64, 64, 3, 3, 0),
    (2, 12, 12, 64, 192, 5, 5, 2),
    (16, 14, 14, 256, 256, 3, 3, 1),
]


@pytest.mark.parametrize("shape", IMPLICIT_SHAPES)
def test_conv_fprop_implicit(shape):
    from maggy_b200.ops import conv

    N, H, W, C, K, R, S, pad = shape
    torch.manual_seed(0)
    x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(K, R, S, C, device="cuda") / (R * S * C) ** 0.5).to(torch.bfloat16)
    y = conv.conv2d_fprop(x, w, pad)
    ref = _ref_conv(x, w, 1, pad)
    assert y.shape == ref.shape
    assert _rel(y, ref) < 1e-2
    yr = conv.conv2d_fprop(x, w, pad, relu=True)
    assert _rel(yr, ref.clamp_min(0)) < 1e-2


@pytest.mark.parametrize("shape", IMPLICIT_SHAPES)
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_conv_wgrad_implicit(shape, splits):
    from maggy_b200.ops import conv

    N, H, W, C, K, R, S, pad = shape
    torch.manual_seed(1)
    x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    P, Q = H + 2 * pad - R + 1, W + 2 * pad - S + 1
    dy = torch.randn(N, P, Q, K, device="cuda").to(torch.bfloat16)
    dw = conv.conv2d_wgrad(x, dy, R, S, pad, splits=splits)
    wref = torch.zeros(K, R, S, C, device="cuda", requires_grad=True)
    _ref_conv(x, wref, 1, pad).backward(dy.float())
    assert _rel(dw, wref.grad) < 1e-2


@pytest.mark.parametrize("cfg", [
    # N, H, W, C, Kout, k, stride, pad
    (4, 14, 14, 64, 64, 3, 1, 1),      # implicit
    (2, 28, 28, 128, 128, 3, 1, 1),    # implicit
    (4, 14, 14, 64, 256, 1, 1, 0),     # pointwise
    (4, 16, 16, 64, 128, 3, 2, 1),     # explicit (strided)
    (2, 15, 15, 32, 64, 1, 2, 0),      # explicit 1x1 / 2
    (2, 32, 32, 3, 64, 7, 2, 3),       # explicit stem (C = 3)
    (2, 10, 10, 16, 24, 3, 1, 1),      # explicit (thin channels)
    (2, 80, 72, 3, 64, 7, 2, 3),       # stem, Q = 36: one full and one partial tile of the staged thin-channel gather
    (1, 20, 70, 5, 16, 3, 1, 1),       # C = 5, stride 1, three tiles per output row
])
def test_conv_module_fwd_bwd(cfg):
    from maggy_b200.ops.conv import Conv2d

    N, H, W, C, K, k, stride, pad = cfg
    torch.manual_seed(2)
    m = Conv2d(C, K, k, stride=stride, padding=pad).cuda()
    ref = torch.nn.Conv2d(C, K, k, stride=stride, padding=pad, bias=False).cuda()
    with torch.no_grad():
        ref.weight.copy_(m.torch_weight().float())
    x = torch.randn(N, C, H, W, device="cuda").to(torch.bfloat16).to(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    x2 = x.float().clone().requires_grad_(True)
    y1 = m(x1)
    y2 = ref(x2)
    assert y1.shape == y2.shape
    assert _rel(y1, y2) < 1e-2
    g = torch.randn_like(y2)
    y1.backward(g.to(torch.bfloat16))
    y2.backward(g.to(torch.bfloat16).float())
    assert _rel(x1.grad, x2.grad) < 1.5e-2
    gw = m.weight.grad.view(K, k, k, C).permute(0, 3, 1, 2)
    assert _rel(gw, ref.weight.grad) < 1.5e-2


def test_resnet_native_vs_torch():
    """Native convolutions + fused norms / pooling against the cuDNN + ATen model with the same weights. Both run in
    bf16 and a deep stack of batch-norms over a batch of 8 amplifies rounding, so each is judged against the SAME model
    evaluated in fp32: the native path may not be further from it than the stock bf16 path (plus a margin)."""
    from maggy_b200.models.resnet import ResNet

    torch.manual_seed(3)
    native = ResNet((1, 1, 1, 1), 10, width=64, conv_impl="native").cuda().to(memory_format=torch.channels_last)
    ref = ResNet((1, 1, 1, 1), 10, width=64, conv_impl="torch").cuda().to(memory_format=torch.channels_last)
    # same weights in both models
    nsd = dict(native.named_modules())
    for name, mod in ref.named_modules():
        if isinstance(mod, torch.nn.Conv2d):
            with torch.no_grad():
                mod.weight.copy_(nsd[name].torch_weight().float())
        elif isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.Linear)):
            mod.load_state_dict(nsd[name].state_dict())
    x = torch.randn(8, 3, 64, 64, device="cuda").to(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), device="cuda")
    l32 = torch.nn.functional.cross_entropy(ref(x), y)          # fp32 oracle
    l32.backward()
    g32 = {n: p.grad.clone() for n, p in ref.named_parameters()}
    ref.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l1 = torch.nn.functional.cross_entropy(native(x), y)
        l2 = torch.nn.functional.cross_entropy(ref(x), y)
    l1.backward()
    l2.backward()
    assert abs(l1.item() - l32.item()) < 5e-2 * max(1.0, abs(l32.item()))
    g1 = native.layer1[0].conv2.weight.grad.view(64, 3, 3, 64).permute(0, 3, 1, 2)
    e_native = _rel(g1, g32["layer1.0.conv2.weight"])
    e_stock = _rel(ref.layer1[0].conv2.weight.grad, g32["layer1.0.conv2.weight"])
    assert e_native < max(0.05, 1.5 * e_stock), (e_native, e_stock)
    e_native = _rel(native.layer4[0].bn3.weight.grad, g32["layer4.0.bn3.weight"])
    e_stock = _rel(ref.layer4[0].bn3.weight.grad, g32["layer4.0.bn3.weight"])
    assert e_native < max(0.05, 1.5 * e_stock), (e_native, e_stock)
    assert torch.isfinite(native.conv1.weight.grad.float()).all()


def test_resnet_trainstep_graph():
    """Whole ResNet step (native convs, fused norms writing straight into the arena, the ATen classifier head adopted
    into it, fused AdamW) replayed as a CUDA graph."""
    from maggy_b200.models.resnet import ResNet
    from maggy_b200.parallel.trainer import build_trainer

    torch.manual_seed(4)
    dev = torch.device("cuda", 0)
    model = ResNet((1, 1, 1, 1), 10, width=64).to(dev).to(memory_format=torch.channels_last)

    def loss_fn(m, x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return torch.nn.functional.cross_entropy(m(x), y)

    flat, opt, step = build_trainer(model, (16, 3, 64, 64), dev, target_shape=(16,),
                                    memory_format=torch.channels_last, loss_fn=loss_fn, lr=2e-3, warmup_iters=2)
    x = torch.randn(16, 3, 64, 64, device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (16,), device=dev)
    losses = [float(step(x, y)) for _ in range(30)]
    assert all(l == l for l in losses)
    assert losses[-1] < losses[0] * 0.7, losses
    assert step.graph is not None and len(flat._autograd_ranges) >= 1
    fc = model.fc.weight   # nn.Linear: gradient produced by ATen autograd, adopted into the arena
    assert fc.grad is not None and fc.grad.data_ptr() == fc.main_grad.data_ptr()
    bn = model.bn1.weight  # fused norm: written by the kernel into the arena, no autograd .grad
    assert bn.grad is None and float(bn.main_grad.float().abs().sum()) > 0


@pytest.mark.parametrize("k,stride,pad", [(1, 1, 0), (3, 1, 1), (1, 2, 0)])
def test_conv_with_skip_adds_the_skip_gradient_in_backward(k, stride, pad):
    """Conv2d(x, with_skip=True) -> (y, x_skip): the gradient arriving through x_skip is added to dgrad (in the GEMM
    epilogue for 1x1 convolutions) — same x.grad as the unfused graph where autograd adds the two."""
    from maggy_b200.ops.conv import Conv2d

    torch.manual_seed(5)
    m = Conv2d(64, 128, k, stride=stride, padding=pad).cuda()
    x = torch.randn(4, 64, 16, 16, device="cuda").to(torch.bfloat16).to(memory_format=torch.channels_last)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y1, xs = m(x1, with_skip=True)
    gy, gs = torch.randn_like(y1), torch.randn_like(x)
    (y1 * gy).sum().backward(retain_graph=True)
    only_conv = x1.grad.clone()
    x1.grad = None
    m.weight.grad = None
    ((y1 * gy).sum() + (xs * gs).sum()).backward()
    y2 = m(x2)
    ((y2 * gy).sum() + (x2 * gs).sum()).backward()
    assert torch.equal(y1, y2)
    assert _rel(x1.grad, x2.grad) < 1e-2
    assert _rel(x1.grad.float() - only_conv.float(), gs) < 2e-2
