"""ctypes loader for the in-tree native libraries.

``cuda()`` returns ``libmaggy_b200_cuda.so`` (all sm_100a kernels). On a machine with a GPU a missing
library is a hard error — the ops never fall back to eager PyTorch silently (the round-end harness
records which in-tree ``.so`` files were actually loaded). ``host()`` returns the CPU-only runtime
(shared-memory control rings, worker launcher) and works everywhere.
"""
from __future__ import annotations

import ctypes
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent.parent
_LIBDIR = _PKG / "lib"
_lock = threading.Lock()
_cuda = None
_host = None

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class NativeLibraryError(RuntimeError):
    pass


def _load(name: str, build_if_missing: bool = True):
    path = _LIBDIR / name
    if not path.exists() and build_if_missing:
        from .. import build as _build

        _build.build(verbose=False)
    if not path.exists():
        raise NativeLibraryError(
            f"{path} is missing: run `python -m maggy_b200.build` (needs nvcc) before using maggy_b200 ops"
        )
    return ctypes.CDLL(str(path))


def _sig(fn, argtypes, restype=c_int):
    fn.argtypes = argtypes
    fn.restype = restype


def cuda():
    """The CUDA kernel library (raises NativeLibraryError if it cannot be loaded)."""
    global _cuda
    if _cuda is not None:
        return _cuda
    with _lock:
        if _cuda is not None:
            return _cuda
        lib = _load("libmaggy_b200_cuda.so")
        P = c_void_p
        _sig(lib.mg_gemm_bf16, [P, P, P, P, P, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P])
        _sig(lib.mg_gemm_auto_splits, [c_int, c_int, c_int, c_int])
        _sig(lib.mg_gemm_fp8, [P, P, P, P, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_float, P, P, c_int, c_int, P])
        _sig(lib.mg_quant_e4m3, [P, P, c_ll, P, P, P])
        _sig(lib.mg_mxfp8_sf_bytes, [c_int, c_int], ctypes.c_ulonglong)
        _sig(lib.mg_quant_mxfp8, [P, c_ll, P, c_ll, P, c_int, c_int, P])
        _sig(lib.mg_gemm_mxfp8, [P, P, P, P, P, P, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_int, P])
        _sig(lib.mg_attn_fwd, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P])
        _sig(lib.mg_attn_bwd, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P])
        _sig(lib.mg_set_device, [c_int])
        _sig(lib.mg_colsum_finalize, [P, c_int, c_int, P, c_int, P])
        _sig(lib.mg_norm_fwd, [P, P, P, P, P, P, c_int, c_int, c_float, c_int, P])
        _sig(lib.mg_norm_bwd_grid, [c_int])
        _sig(lib.mg_norm_bwd, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P])
        _sig(lib.mg_norm_bwd_res, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P])
        _sig(lib.mg_norm_bwd_col, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P])
        _sig(lib.mg_norm_bwd_supports_col, [c_int])
        _sig(lib.mg_adam_flat, [P, P, P, P, P, c_ll, c_ll, c_float, c_float, c_float, c_float,
                                c_float, c_int, c_float, P, c_float, P, P, P])
        _sig(lib.mg_sumsq_bf16, [P, c_ll, P, P])
        _sig(lib.mg_sumsq_f32, [P, c_ll, P, P])
        _sig(lib.mg_adam_flat_f32, [P, P, P, P, c_ll, c_ll, c_float, c_float, c_float, c_float, c_float, c_int, c_float,
                                    P, c_float, P, P, P])
        _sig(lib.mg_cast_f32_bf16, [P, P, c_ll, P])
        _sig(lib.mg_ce_fwd_bwd, [P, P, P, c_int, c_int, c_int, c_ll, c_float, P, c_int, P])
        _sig(lib.mg_colsum_slabs, [c_int])
        _sig(lib.mg_colsum_bf16, [P, c_int, c_int, c_ll, P, P, c_int, P])
        _sig(lib.mg_gelu_fwd, [P, P, c_ll, P])
        _sig(lib.mg_gelu_bwd, [P, P, P, c_ll, P])
        _sig(lib.mg_embed_fwd, [P, P, P, P, c_int, c_int, c_int, P])
        _sig(lib.mg_embed_bwd, [P, P, P, P, c_int, c_int, c_int, c_int, P])
        for name, sig in _OPTIONAL_CUDA.items():
            if hasattr(lib, name):
                _sig(getattr(lib, name), *sig)
        _cuda = lib
        return lib


# symbols added by later kernels; registered here so older builds still load
_OPTIONAL_CUDA: dict = {}


def register_cuda_symbol(name: str, argtypes, restype=c_int):
    _OPTIONAL_CUDA[name] = (argtypes, restype)
    if _cuda is not None and hasattr(_cuda, name):
        _sig(getattr(_cuda, name), argtypes, restype)


def host():
    """The host-only runtime library (rings, launcher)."""
    global _host
    if _host is not None:
        return _host
    with _lock:
        if _host is None:
            _host = _load("libmaggy_b200_host.so")
        return _host


def check(rc: int, what: str):
    if rc != 0:
        raise NativeLibraryError(f"{what} failed with code {rc}")
