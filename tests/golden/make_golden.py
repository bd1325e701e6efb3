"""Generate tests/golden/ekf_golden.npz from oracle/_ref (the unmodified reference sources compiled here).

Run in the container that has /root/reference:  python tests/golden/make_golden.py
The fixtures are small (a few hundred KB) and committed; the GPU box and later rounds check the oracle
restatement and the CUDA path against them without needing /root/reference.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from shermbot_navigation_b200 import synth  # noqa: E402


def main():
    oracle.build()
    ref = oracle.load("ref")
    assert ref.flavour == "reference"
    sc = synth.ekf_scenario(4, 25, n=12, seed=2024)
    out = dict(n=12, robot0=sc["robot0"], map0=sc["map0"], Q=sc["Q"], R=sc["R"], twists=sc["twists"], z=sc["z"], ids=sc["ids"])
    for tag, ids in (("known", sc["ids"]), ("unknown", None)):
        r = ref.ekf_run(12, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"], sc["z"], ids, trace=True)
        for k in ("x", "sigma", "seen", "ids_out", "trace"):
            out[f"{tag}_{k}"] = r[k]
        # the state after the first step (every landmark's first touch lies behind it): the warm start from which a GPU path can be
        # held to 1e-9 without sharing the reference's libm (gate L1)
        r1 = ref.ekf_run(12, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][:1], sc["z"][:1], None if ids is None else ids[:1])
        for k in ("x", "sigma", "seen"):
            out[f"{tag}_{k}1"] = r1[k]
    s = synth.scan_scenario(64, seed=4242, noise_sigma=0.001)
    sd = ref.scan_detect_batch(s["ranges"], s["min_range"], s["max_range"])
    out["ranges"] = s["ranges"]
    out["scan_cluster_of_beam"] = sd["cluster_of_beam"]
    out["scan_n_clusters"] = sd["n_clusters"]
    out["scan_n_circles"] = sd["n_circles"]
    out["scan_circles"] = sd["circles"]
    p = Path(__file__).parent / "ekf_golden.npz"
    np.savez_compressed(p, **out)
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
