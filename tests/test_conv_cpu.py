"""CPU-side checks of the conv / ResNet front-end (the kernels themselves are covered by tests/test_gpu_conv.py)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_conv2d_module_cpu_path_matches_torch():
    from maggy_b200.ops.conv import Conv2d

    torch.manual_seed(0)
    for (cin, cout, k, stride, pad) in [(16, 32, 3, 1, 1), (8, 8, 1, 2, 0), (3, 16, 7, 2, 3)]:
        ours = Conv2d(cin, cout, k, stride=stride, padding=pad, dtype=torch.float32)
        ref = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)
        ours.load_torch_weight(ref.weight)
        assert ours.weight.shape == (cout, k * k * cin)              # KRSC flattened: the GEMM operand
        assert torch.equal(ours.torch_weight(), ref.weight)
        x = torch.randn(2, cin, 12, 12)
        assert torch.allclose(ours(x), ref(x), atol=1e-5)


def test_conv2d_weight_layout_survives_channels_last_conversion():
    from maggy_b200.ops.conv import Conv2d

    m = torch.nn.Sequential(Conv2d(8, 8, 3, padding=1)).to(memory_format=torch.channels_last)
    assert m[0].weight.is_contiguous() and m[0].weight.dim() == 2


def test_resnet_cpu_forward_backward_both_conv_impls():
    from maggy_b200.models.resnet import resnet18ish

    x = torch.randn(2, 3, 32, 32)
    for impl in ("native", "torch"):
        m = resnet18ish(10, conv_impl=impl)
        y = m(x)
        assert y.shape == (2, 10)
        y.float().sum().backward()
        assert all(p.grad is not None for p in m.parameters())


def test_implicit_gemm_eligibility():
    from maggy_b200.ops.conv import implicit_ok

    assert implicit_ok(64, 64, 1) and implicit_ok(256, 512, 1)
    assert not implicit_ok(3, 64, 1) and not implicit_ok(64, 64, 2) and not implicit_ok(64, 96, 1)


def test_bench_reference_arm_contract_without_gpu():
    """`bench.py --impl reference` must print one JSON line with "impl": "reference" and exit 0 even when it cannot
    measure (no GPU in the build container)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["impl"] == "reference" and ("unavailable" in out or "value" in out)


def test_reference_imports_with_the_stand_ins():
    import pytest

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "maggy")):
        pytest.skip("baseline/_ref not installed")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import refenv; refenv.install()\n"
        "from maggy.experiment import experiment\n"
        "from maggy.config import BaseConfig\n"
        "import os, tempfile; os.chdir(tempfile.mkdtemp())\n"
        "r = experiment.lagom(lambda: {'metric': 1.0}, BaseConfig(name='standin_probe'))\n"
        "print('REF_OK', r)\n" % (os.path.join(ROOT, "baseline"), ref_dir))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "REF_OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_fused_batchnorm_module_falls_back_on_cpu():
    """ops.batchnorm.BatchNorm2d on CPU tensors = relu(batch_norm(x) + residual) through ATen (same state_dict keys)."""
    import torch

    from maggy_b200.ops.batchnorm import BatchNorm2d

    torch.manual_seed(0)
    bn, ref = BatchNorm2d(8, relu=True), torch.nn.BatchNorm2d(8)
    assert set(bn.state_dict()) == set(ref.state_dict())
    x, r = torch.randn(4, 8, 5, 5), torch.randn(4, 8, 5, 5)
    assert torch.allclose(bn(x, r), torch.relu(ref(x) + r), atol=1e-6)
    assert torch.allclose(bn.running_var, ref.running_var)


def test_native_pool_and_skip_conv_fall_back_on_cpu():
    """ops.pool.MaxPool2d and Conv2d(..., with_skip=True) on CPU tensors: plain ATen results, same call signatures as on
    the GPU (ResNet bottlenecks use both on every device)."""
    import torch

    from maggy_b200.models.resnet import resnet18ish
    from maggy_b200.ops.conv import Conv2d
    from maggy_b200.ops.pool import MaxPool2d

    torch.manual_seed(0)
    x = torch.randn(2, 8, 9, 9)
    assert torch.equal(MaxPool2d(3, stride=2, padding=1)(x), torch.nn.MaxPool2d(3, stride=2, padding=1)(x))
    conv = Conv2d(8, 16, 1, dtype=torch.float32)
    y, skip = conv(x, with_skip=True)
    assert skip is x and torch.equal(y, conv(x))
    net = resnet18ish(10)
    out = net(torch.randn(2, 3, 32, 32))
    out.sum().backward()
    assert out.shape == (2, 10) and net.conv1.weight.grad is not None
