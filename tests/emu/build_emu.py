"""Build ``tests/emu/libsnrf_emu.so`` (host emulation of the simple kernels - TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsnrf_emu.so")
CSRC = os.path.normpath(os.path.join(HERE, "..", "..", "segment-anything-in-nerf_b200", "csrc"))


def build(force: bool = False) -> str:
    deps = [os.path.join(HERE, "emu.cu")] + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # host side only matters; -fmad=false / -ffp-contract=off keep a*b+c as two roundings like torch on the CPU
    cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-Xcompiler",
           "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, os.path.join(HERE, "emu.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc (emu) failed:\n" + proc.stdout + proc.stderr)
    return LIB


def load() -> ctypes.CDLL:
    """Build if needed and load.  Without nvcc (and without a prebuilt library) the calling test is skipped: the
    emulation is a checker for the build container, not a requirement of the package."""
    import shutil

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(LIB) and not (os.path.exists(nvcc) or shutil.which("nvcc")):
        import pytest

        pytest.skip("nvcc not available: cannot build the host emulation of the kernel bodies")
    return ctypes.CDLL(build())


if __name__ == "__main__":
    print(build(force=True))
