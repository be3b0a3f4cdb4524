"""Compile-only guards (no GPU needed): every .cu builds for sm_100a and the SASS contains the instructions
that prove the Blackwell-native path — a cheap guard against silently falling back."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "maggy_b200", "lib", "libmaggy_b200_cuda.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("nvcc") is None or shutil.which("cuobjdump") is None:
        pytest.skip("CUDA toolchain not available")
    from maggy_b200 import build

    build.build(verbose=False)
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = []
        elif cur is not None:
            per[cur].append(line)
    return {k: "\n".join(v) for k, v in per.items()}


def _kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_only_sm100a_code_is_built():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_gemm_uses_tcgen05_tmem_and_tma(sass):
    ks = _kernels(sass, "gemm_bf16_tcgen05_kernel")
    assert len(ks) >= 8                       # block_n x operand-major instantiations
    fp8 = {k for k in ks if "Lb1EEEv" in k}    # <..., FP8 = true> instantiations
    assert len(fp8) >= 2
    for name, text in ks.items():
        if name in fp8:
            assert "UTCQMMA" in text, name    # tcgen05.mma kind::f8f6f4
        else:
            assert "UTCHMMA" in text, name    # tcgen05.mma kind::f16
        assert "LDTM" in text, name           # tcgen05.ld (TMEM -> registers)
        assert "UTMALDG" in text, name        # TMA loads
        assert "UTMASTG" in text, name        # TMA store epilogue
        assert "SYNCS" in text, name          # mbarrier pipeline
        assert "HMMA." not in text.replace("UTCHMMA", ""), name   # no legacy mma.sync path


def test_pair_gemm_uses_cta_group_2(sass):
    ks = {k: v for k, v in _kernels(sass, "gemm_bf16_tcgen05_kernel").items() if "ELi2ELb" in k}   # MODE = 2
    assert len(ks) >= 5
    for name, text in ks.items():
        assert re.search(r"UTC[HQ]MMA\.2CTA", text), name       # one MMA spanning both SMs of the pair
        assert "UTMALDG.2D.2CTA" in text, name                  # TMA loads signalling the leader's mbarrier
        assert "UTCBAR.2CTA.MULTICAST" in text, name            # commit releases both CTAs' smem slots


def test_conv_kernels_are_implicit_gemm_on_tcgen05(sass):
    fprop = _kernels(sass, "conv2d_fprop_tcgen05_kernel")
    wgrad = _kernels(sass, "conv2d_wgrad_tcgen05_kernel")
    assert len(fprop) == 3 and len(wgrad) == 3
    for name, text in {**fprop, **wgrad}.items():
        assert "UTCHMMA" in text and "LDTM" in text, name
        assert "UTMALDG.4D" in text, name                       # shifted pixel boxes straight from NHWC
    for name, text in fprop.items():
        assert "UTMASTG.4D" in text, name                       # output pixel box stored by TMA
    for name, text in wgrad.items():
        assert "REDG.E.ADD.F32x4" in text, name                 # split-pixel partial sums


def test_allgather_gemm_mixes_peer_loads_with_tensor_core_tiles(sass):
    (text,) = _kernels(sass, "ag_gemm_kernel").values()
    assert "UTCHMMA" in text and "UTMALDG" in text and "UTMASTG" in text
    assert "UBLKCP.S.G" in text and "UBLKCP.G.S" in text          # bulk copies: peer HBM -> smem -> local HBM
    assert re.search(r"(LDG|STG)\.E(\.\d+)?\.STRONG\.SYS", text)    # cross-GPU ready / done flags


def test_fused_reduce_scatter_adam_uses_multimem_and_sys_scope(sass):
    ks = _kernels(sass, "rs_adam_kernel")
    assert len(ks) == 2
    joined = "\n".join(ks.values())
    assert "LDGMC" in joined                                        # multimem.ld_reduce (NVLS in-switch add)
    assert re.search(r"LDG\.E\.128\.STRONG\.SYS", joined)          # peer loads of the P2P variant
    assert re.search(r"STG\.E(\.128)?\.STRONG\.SYS", joined)       # release flags / multicast stores


def test_control_plane_kernels_publish_at_system_scope(sass):
    (rep,) = _kernels(sass, "report_store_kernel").values()
    (sch,) = _kernels(sass, "sched_kernel").values()
    assert re.search(r"STG\.E\.64\.STRONG\.SYS", rep)
    assert re.search(r"LDG\.E\.64\.STRONG\.SYS", sch) and "NANOSLEEP" in sch


def test_no_kernel_spills_except_wide_norm_variants():
    logdir = os.path.join(ROOT, "maggy_b200", "lib", "obj")
    bad = []
    for f in os.listdir(logdir):
        if not f.endswith(".ptxas.log"):
            continue
        fn = None
        for line in open(os.path.join(logdir, f)):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = m.group(1)
            m = re.search(r"(\d+) bytes spill stores", line)
            if m and int(m.group(1)) > 64 and fn and "norm_" not in fn:
                bad.append((f, fn, int(m.group(1))))
    assert not bad, bad
