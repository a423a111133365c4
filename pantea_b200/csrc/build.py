"""Builds libpantea_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pantea_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SOURCES = ["potential.cu", "neighbor.cu", "acsf.cu", "acsf2.cu", "md.cu", "microbench.cu", "lj.cu", "scaler.cu", "halo.cu", "mgpu.cu"]
HEADERS = [HERE / "internal.cuh", HERE / "math.cuh", HERE / "acsf_common.cuh", ROOT / "include" / "pantea_b200.h"]
LIB = HERE.parent / "libpantea_b200.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in [*(HERE / s for s in SOURCES), *HEADERS])
    return LIB.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    common = [nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", f"-I{ROOT / 'include'}", f"-I{HERE}"]
    common += os.environ.get("PANTEA_DEFINES", "").split()
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        procs.append((src, subprocess.Popen(common + ["-c", str(HERE / src), "-o", str(obj)],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0 or verbose:
            print(f"--- {src}\n{out}", file=sys.stderr)
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, *ARCH, "-shared", "-o", str(LIB), *[str(objdir / (s + ".o")) for s in SOURCES]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr, file=sys.stderr)
        raise RuntimeError("link failed")
    return LIB


def build_ffi_shim() -> Path:
    """Compile csrc/xla_ffi_shim.cc (XLA FFI handlers over the C ABI) into libpantea_b200_ffi.so.  Needs jaxlib's FFI
    headers; raises RuntimeError with the reason when they are not available (this image: no jax)."""
    try:
        from jax import ffi as jax_ffi
    except ImportError as exc:
        raise RuntimeError(f"XLA FFI shim not built: jax.ffi is not importable here ({exc})") from exc
    lib = build()
    out = HERE.parent / "libpantea_b200_ffi.so"
    cmd = [_nvcc(), *ARCH, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", f"-I{ROOT / 'include'}",
           f"-I{jax_ffi.include_dir()}", str(HERE / "xla_ffi_shim.cc"), "-o", str(out), f"-L{lib.parent}",
           "-lpantea_b200", f"-Xlinker=-rpath={lib.parent}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr, file=sys.stderr)
        raise RuntimeError("XLA FFI shim: compilation failed")
    return out


if __name__ == "__main__":
    if "--ffi" in sys.argv:
        print(build_ffi_shim())
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
