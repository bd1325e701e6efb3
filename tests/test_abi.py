"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/nuslam_b200.h declares, and refuses loudly to compute without a GPU (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "nuslam_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nuslam_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cuda_lib):
    import ctypes
    lib = cuda_lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nuslam_b200.h but not exported"
    assert sorted(cuda_lib.EXPORTS) == syms
    assert lib.nuslam_version() == 101


def test_library_is_sm100a_only(cuda_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", str(cuda_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback(cuda_lib):
    """Without a device the engine must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cuda_lib.NuslamError):
        cuda_lib.BatchedExtendedKalman(np.zeros((2, 3)), n_landmarks=12)
    with pytest.raises(cuda_lib.NuslamError):
        cuda_lib.cartesian2polar(np.ones((4, 2)))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = ROOT / "shermbot-navigation_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list((ROOT / "include").rglob("*")):
        if p.is_file():
            t = p.read_text(errors="ignore")
            assert "import oracle" not in t and "from oracle" not in t and "libnuslam_oracle" not in t and "libnuslam_ref" not in t, p


def test_default_config(cuda_lib):
    import ctypes as C
    cfg = cuda_lib.EkfConfig()
    cuda_lib.lib().nuslam_ekf_default_config(C.byref(cfg), 12)
    assert cfg.n_landmarks == 12 and cfg.mode == 0
    assert list(cfg.Q) == [0.1, 0, 0, 0, 0.1, 0, 0, 0, 0.1] and list(cfg.R) == [0.001, 0, 0, 0.001]
    assert cfg.assoc_min == 0.01 and cfg.assoc_max == 60


def build_abi_main(tmp_path, cuda_lib):
    import subprocess
    exe = tmp_path / "abi_main"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "abi_main.c"), "-o", str(exe),
           str(cuda_lib.LIB_PATH), f"-Wl,-rpath,{cuda_lib.LIB_PATH.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_every_entry_point_links(cuda_lib, tmp_path):
    """The boundary is a C ABI: include/nuslam_b200.h compiles as C99 (-pedantic -Werror), a C program links every declared entry point,
    and the calls that need no device behave (version, default config, refusal of a null config)."""
    import subprocess
    exe = build_abi_main(tmp_path, cuda_lib)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    assert out["VERSION"] == "101" and int(out["ENTRY_POINTS"]) == len(declared_symbols())
    assert "n=12 mode=0 Q00=0.1 R00=0.001 amin=0.01 amax=60 options=0 prior=2147483647.0" in out["DEFAULT"]
    assert "null config" in out["NULLCFG"]


@pytest.mark.gpu
def test_c_program_runs_the_batched_loop(cuda_lib, tmp_path):
    """INTEGRATION.md section 3 from plain C: create, init, three fused steps with host buffers, get_state."""
    import subprocess
    exe = build_abi_main(tmp_path, cuda_lib)
    r = subprocess.run([str(exe), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    state = [ln for ln in r.stdout.splitlines() if ln.startswith("STATE")][0].split()
    assert state[-1] == "0" and abs(float(state[1]) - 0.06) < 0.05
