"""Build the in-tree CUDA library (sm_100a only). nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB = HERE / "libnuslam_b200.so"
SOURCES = [HERE / "csrc" / "nuslam_b200.cu"]
HEADERS = sorted((HERE / "csrc").glob("*.cuh")) + [ROOT / "include" / "nuslam_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{HERE / 'csrc'}", "-o", str(LIB), *map(str, SOURCES)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libnuslam_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
