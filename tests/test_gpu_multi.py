"""Multi-GPU checks (need >= 2 devices; skipped otherwise): fused reduce-scatter+Adam+all-gather kernel."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_rs_adam_matches_nccl_oracle():
    import torch

    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi", "rs_adam_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "RS_ADAM_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_allgather_fused_gemm():
    import torch

    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "multi", "ag_gemm_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "AG_GEMM_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_sharded_linear_trains_like_the_unsharded_model():
    import torch

    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29521",
           os.path.join(ROOT, "tests", "multi", "sharded_linear_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED_LINEAR_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_block_scaled_fp8_fsdp_path():
    """K3-MX (all-gather of MXFP8 shards fused into the block-scaled GEMM), the peer-memory reduce-scatter kernel and a
    ShardedLinearMX training step against PyTorch / NCCL oracles."""
    import torch

    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29523", os.path.join(ROOT, "tests", "multi", "mx_fsdp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "MX_FSDP_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_generic_module_takes_the_fused_data_parallel_path():
    """A plain fp32 nn.Sequential under the class-factory wrapper + scoped optimizer patch: K2 broadcast, fp32-arena
    K1 (AdamW at zero_lvl 0 and 1) and the arena all-reduce hook (SGD) against DDP oracles."""
    import torch

    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29525", os.path.join(ROOT, "tests", "multi", "generic_dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "GENERIC_DP_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
