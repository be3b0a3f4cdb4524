"""In-tree native build for maggy_b200 (sm_100a only).

Compiles every ``csrc/**/*.cu`` / ``*.cpp`` into two shared objects next to the package:

* ``maggy_b200/lib/libmaggy_b200_cuda.so`` — all device kernels + their host launchers (nvcc,
  ``-gencode arch=compute_100a,code=sm_100a -lineinfo``). Loaded with ctypes; needs a CUDA driver.
* ``maggy_b200/lib/libmaggy_b200_host.so`` — the host-only runtime (shared-memory control rings,
  worker-pool launcher). Plain g++, loads on CPU-only machines.

The objects are cached by content hash, so repeated ``build()`` calls are cheap. The built ``.so``
files are git-ignored but travel to the GPU box with the ``gpurun`` snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "lib" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
CXX_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-pthread"]

CUDA_LIB = LIBDIR / "libmaggy_b200_cuda.so"
HOST_LIB = LIBDIR / "libmaggy_b200_host.so"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(path: Path, flags: list[str]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    h.update(path.read_bytes())
    # headers are few: hash them all so an edit to a .cuh invalidates every object
    for hdr in sorted(CSRC.rglob("*.cuh")) + sorted(CSRC.rglob("*.h")):
        h.update(hdr.read_bytes())
    return h.hexdigest()[:16]


def _compile_one(src: Path, cuda: bool, verbose: bool) -> Path:
    flags = NVCC_FLAGS if cuda else CXX_FLAGS
    tag = _digest(src, flags)
    obj = OBJDIR / f"{src.stem}.{tag}.o"
    if obj.exists():
        return obj
    for old in OBJDIR.glob(f"{src.stem}.*.o"):
        old.unlink()
    if cuda:
        cmd = [_nvcc(), *flags, "-I", str(CSRC / "include"), "-c", str(src), "-o", str(obj)]
    else:
        cmd = [os.environ.get("CXX", "g++"), *flags, "-I", str(CSRC / "include"), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    log = OBJDIR / f"{src.stem}.ptxas.log"
    log.write_text(res.stderr)
    if verbose:
        spills = [l for l in res.stderr.splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
        print(f"[maggy_b200.build] {src.relative_to(ROOT)} -> {obj.name}" + (f"  ({len(spills)} kernels spill)" if spills else ""))
    return obj


def build(verbose: bool = True, force: bool = False) -> dict:
    """Compile all native code in-tree. Returns {"cuda": path, "host": path}."""
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    if force:
        for f in OBJDIR.glob("*.o"):
            f.unlink()
    cu_srcs = sorted(CSRC.rglob("*.cu"))
    cpp_srcs = sorted(CSRC.rglob("*.cpp"))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        cu_objs = list(ex.map(lambda s: _compile_one(s, True, verbose), cu_srcs))
        cpp_objs = list(ex.map(lambda s: _compile_one(s, False, verbose), cpp_srcs))

    def _link(out: Path, objs: list[Path], cmd_prefix: list[str]):
        stamp = hashlib.sha256(" ".join(sorted(o.name for o in objs)).encode()).hexdigest()[:16]
        stamp_file = OBJDIR / (out.name + ".stamp")
        if out.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
            return
        cmd = [*cmd_prefix, "-o", str(out), *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        stamp_file.write_text(stamp)
        if verbose:
            print(f"[maggy_b200.build] linked {out.relative_to(ROOT)}")

    if cu_objs:
        _link(CUDA_LIB, cu_objs, [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                                  "-Xcompiler", "-fPIC", "-lpthread"])
    if cpp_objs:
        _link(HOST_LIB, cpp_objs, [os.environ.get("CXX", "g++"), "-shared", "-pthread", "-lrt"])
    return {"cuda": str(CUDA_LIB), "host": str(HOST_LIB)}


if __name__ == "__main__":
    out = build(verbose=True, force="--force" in sys.argv)
    print(out)
