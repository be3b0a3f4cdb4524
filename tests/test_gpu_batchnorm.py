"""Fused NHWC BatchNorm (+ReLU, +residual) kernels (csrc/ops/batchnorm.cu) against a plain fp32 PyTorch reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, r):
    a, r = a.float(), r.float()
    return ((a - r).norm() / (r.norm() + 1e-12)).item()


@pytest.mark.parametrize("shape", [(8, 64, 56, 56), (4, 256, 14, 14), (2, 2048, 7, 7), (3, 24, 5, 5), (16, 128, 28, 28)])
@pytest.mark.parametrize("relu,res", [(False, False), (True, False), (True, True), (False, True)])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
def test_batchnorm_matches_fp32_reference(shape, relu, res, pdtype):
    from maggy_b200.ops.batchnorm import BatchNorm2d

    torch.manual_seed(7)
    N, C, H, W = shape
    dev = "cuda"
    x = (torch.randn(shape, device=dev) * 1.7 + 0.4).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    g = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bn = BatchNorm2d(C, relu=relu).to(dev)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(C, device=dev))
        bn.bias.copy_(0.2 * torch.randn(C, device=dev))
    bn.weight.data = bn.weight.data.to(pdtype)
    bn.bias.data = bn.bias.data.to(pdtype)
    ref = torch.nn.BatchNorm2d(C).to(dev)
    with torch.no_grad():
        ref.weight.copy_(bn.weight.float())
        ref.bias.copy_(bn.bias.float())

    x1 = x.clone().requires_grad_(True)
    r1 = r.clone().requires_grad_(True) if res else None
    y = bn(x1, r1)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    y.backward(g)

    x2 = x.float().requires_grad_(True)
    r2 = r.float().requires_grad_(True) if res else None
    o = ref(x2)
    if res:
        o = o + r2
    if relu:
        o = torch.relu(o)
    o.backward(g.float())

    assert _rel(y, o) < 1e-2
    assert _rel(x1.grad, x2.grad) < 2e-2
    if res:
        assert _rel(r1.grad, r2.grad) < 1e-2
    assert _rel(bn.weight.grad, ref.weight.grad) < 2e-2
    assert _rel(bn.bias.grad, ref.bias.grad) < 2e-2
    assert _rel(bn.running_mean, ref.running_mean) < 1e-2
    assert _rel(bn.running_var, ref.running_var) < 1e-2
    assert int(bn.num_batches_tracked) == 1

    # frozen statistics
    bn.eval(), ref.eval()
    with torch.no_grad():
        ye = bn(x, r)
        oe = ref(x.float()) + (r.float() if res else 0)
        oe = torch.relu(oe) if relu else oe
    assert _rel(ye, oe) < 1e-2


def test_batchnorm_main_grad_convention_accumulates():
    from maggy_b200.ops.batchnorm import BatchNorm2d

    torch.manual_seed(1)
    bn = BatchNorm2d(64, relu=True).cuda()
    x = torch.randn(4, 64, 8, 8, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for p in (bn.weight, bn.bias):
        p.main_grad = torch.zeros_like(p)
        p._mg_written = False
    xg = x.clone().requires_grad_(True)
    bn(xg).sum().backward()
    first = bn.weight.main_grad.clone(), bn.bias.main_grad.clone()
    assert bn.weight.grad is None and first[1].abs().sum() > 0
    bn(x.clone().requires_grad_(True)).sum().backward()   # _mg_written is set: accumulates
    assert _rel(bn.bias.main_grad, 2 * first[1]) < 1e-3


def test_resnet_step_with_fused_norm_trains():
    """ResNet bottleneck stack on the native convolutions + fused norms: loss goes down, no ATen batch-norm kernel runs."""
    from maggy_b200.models.resnet import ResNet
    from maggy_b200.ops.batchnorm import BatchNorm2d

    torch.manual_seed(0)
    model = ResNet((1, 1, 1, 1), num_classes=10, width=64).cuda().to(memory_format=torch.channels_last)
    assert all(isinstance(m, BatchNorm2d) for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d))
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    x = torch.randn(16, 3, 64, 64, device="cuda").to(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (16,), device="cuda")
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] * 0.7, losses


@pytest.mark.parametrize("shape,k,s,p", [((4, 64, 32, 32), 3, 2, 1), ((2, 16, 15, 17), 3, 2, 1), ((2, 8, 9, 9), 2, 2, 0),
                                         ((1, 24, 12, 12), 3, 1, 1)])
def test_maxpool_matches_aten_including_relu_ties(shape, k, s, p):
    """Post-ReLU inputs are full of exact ties (zeros): values, and the routing of the gradient to the FIRST maximum of
    every window, must match ATen bit for bit."""
    from maggy_b200.ops.pool import MaxPool2d

    torch.manual_seed(3)
    x = torch.relu(torch.randn(shape, device="cuda")).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y1 = MaxPool2d(k, stride=s, padding=p)(x1)
    y2 = torch.nn.MaxPool2d(k, stride=s, padding=p)(x2)
    assert y1.is_contiguous(memory_format=torch.channels_last) and torch.equal(y1, y2)
    g = torch.randn_like(y2)
    y1.backward(g)
    y2.backward(g)
    assert _rel(x1.grad, x2.grad) < 4e-3   # overlapping windows: ATen sums in fp32 atomics, here fp32 then one rounding
    assert torch.equal(x1.grad == 0, x2.grad == 0)


@pytest.mark.parametrize("relu,res", [(True, True), (False, False)])
def test_batchnorm_frozen_statistics_backward(relu, res):
    """eval() mode with gradients (frozen-BN fine-tuning): y uses the running statistics, dx = gamma / std * dz and
    dgamma / dbeta are still produced."""
    from maggy_b200.ops.batchnorm import BatchNorm2d

    torch.manual_seed(2)
    shape = (4, 64, 12, 12)
    x = (torch.randn(shape, device="cuda") * 1.3 + 0.2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn(shape, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    g = torch.randn(shape, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bn, ref = BatchNorm2d(64, relu=relu).cuda(), torch.nn.BatchNorm2d(64).cuda()
    with torch.no_grad():
        for m in (bn, ref):
            m.running_mean.copy_(torch.linspace(-0.5, 0.5, 64))
            m.running_var.copy_(torch.linspace(0.5, 2.0, 64))
            m.weight.copy_(torch.linspace(0.5, 1.5, 64))
            m.bias.copy_(torch.linspace(-0.2, 0.2, 64))
    bn.eval(), ref.eval()
    x1 = x.clone().requires_grad_(True)
    r1 = r.clone().requires_grad_(True) if res else None
    bn(x1, r1).backward(g)
    x2 = x.float().requires_grad_(True)
    r2 = r.float().requires_grad_(True) if res else None
    o = ref(x2) + (r2 if res else 0)
    (torch.relu(o) if relu else o).backward(g.float())
    assert _rel(x1.grad, x2.grad) < 1e-2
    assert _rel(bn.weight.grad, ref.weight.grad) < 2e-2 and _rel(bn.bias.grad, ref.bias.grad) < 2e-2
    if res:
        assert _rel(r1.grad, r2.grad) < 1e-2
    assert torch.equal(bn.running_mean, ref.running_mean)      # untouched
