"""Build the in-tree CUDA library (sm_100a only). nvcc cross-compiles without a GPU.

Two translation units (csrc/ekf_fast_api.cuh says why): nuslam_b200.cu is compiled whole-program; ekf_fast_tu.cu -- the FAST filter
kernels, which launch the oracle-order list kernel themselves -- as relocatable device code, device-linked against cudadevrt."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB = HERE / "libnuslam_b200.so"
OBJ = ROOT / "build" / "obj"
SOURCES = [HERE / "csrc" / "nuslam_b200.cu", HERE / "csrc" / "ekf_fast_tu.cu"]
HEADERS = sorted((HERE / "csrc").glob("*.cuh")) + [ROOT / "include" / "nuslam_b200.h"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DNUSLAM_TAIL_LAUNCH=1"]
# per source: extra compile flags (-dc = relocatable device code)
UNIT_FLAGS = {"nuslam_b200.cu": ["-c"], "ekf_fast_tu.cu": ["-dc"]}


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


def _run(cmd, verbose):
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libnuslam_b200.so: " + " ".join(cmd[:1] + cmd[-3:]))


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib: Path = LIB, tag: str = "") -> Path:
    """`extra_flags` / `lib` / `tag`: kernel-experiment builds (tools/build_variants.sh) into another file."""
    if not force and lib == LIB and not stale():
        return LIB
    OBJ.mkdir(parents=True, exist_ok=True)
    inc = [f"-I{ROOT / 'include'}", f"-I{HERE / 'csrc'}"]
    objs = []
    cmds = []
    for src in SOURCES:
        o = OBJ / (src.stem + tag + ".o")
        objs.append(o)
        cmd = [nvcc(), *NVCC_FLAGS, *extra_flags, *inc, *UNIT_FLAGS[src.name], "-o", str(o), str(src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmds.append(cmd)
    with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
        list(ex.map(lambda c: _run(c, verbose), cmds))
    # nvcc device-links the relocatable object (against cudadevrt) and links the shared library
    _run([nvcc(), *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", str(lib), *map(str, objs), "-lcudadevrt"], verbose)
    if tag:   # experiment builds leave no objects behind (build/ travels to the GPU box with every gpurun call)
        for o in objs:
            o.unlink(missing_ok=True)
    return lib


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 3 and sys.argv[1] == "--variant":   # python build.py --variant NAME [flags...]
        out = ROOT / "build" / "variants" / f"lib_{sys.argv[2]}.so"
        out.parent.mkdir(parents=True, exist_ok=True)
        build(force=True, verbose=True, extra_flags=sys.argv[3:], lib=out, tag="_" + sys.argv[2])
        print(out)
    else:
        build(force=True, verbose=True)
        print(LIB)
