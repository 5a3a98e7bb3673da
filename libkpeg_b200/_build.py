"""In-tree build of the native artefacts (nvcc / g++ via make).  Used by __graft_entry__.build()
and, lazily, by the ctypes loader when a library is missing or older than its sources."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "lib"


def _run(cmd, cwd):
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(map(str, cmd))}\n{proc.stdout[-4000:]}")
    return proc.stdout


def build_product(targets=("all",)) -> None:
    """libkpeg_cuda.so (sm_100a kernels + C ABI), libkpeg_synth.so, kpeg CLI."""
    _run(["make", "-j8", *targets], cwd=PKG)


def build_oracle() -> None:
    """oracle/_ref: the CPU restatement, and the unmodified reference when /root/reference exists."""
    _run(["make", "all"], cwd=ROOT / "oracle")
    # the reference's own executable with its hot path replaced by the C ABI, when
    # /root/reference is present; needs libkpeg_cuda.so, so build_product() first
    import sys
    _run([sys.executable, str(ROOT / "oracle" / "hybrid" / "build.py")], cwd=ROOT)


def build_emu() -> None:
    """tests/emu: CPU single-stepper of the kernel logic (test infrastructure)."""
    emu = ROOT / "tests" / "emu"
    out = emu / "libkpeg_emu.so"
    src = emu / "emu.cpp"
    deps = [src, *sorted((PKG / "csrc").glob("*.h"))]
    if out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return
    _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
          "-Wno-unknown-pragmas", f"-I{ROOT / 'include'}", f"-I{PKG / 'csrc'}", "-o", str(out), str(src)], cwd=emu)


def have_nvcc() -> bool:
    from shutil import which
    return which(os.environ.get("NVCC", "nvcc")) is not None
