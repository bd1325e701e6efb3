#!/usr/bin/env python
"""Debug: large-map mode from scratch against the oracle, step by step."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle
from shermbot_navigation_b200 import nuslam, synth
orc = oracle.load("ref") if oracle.available("ref") else oracle.load("port")
n, B, T = 20, 2, 3
sc = synth.ekf_scenario(B, T, n=n, seed=62)
rel = lambda a, b: np.abs(a - b).max() / max(np.abs(a).max(), np.abs(b).max(), 1e-300)
for mode in ("large", "strict"):
    eng = nuslam.BatchedExtendedKalman(sc["robot0"], sc["map0"], sc["Q"], sc["R"], mode=mode)
    for t in range(T):
        eng.step(sc["twists"][t], sc["z"][t], sc["ids"][t])
        r = orc.ekf_run(n, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][:t + 1], sc["z"][:t + 1], sc["ids"][:t + 1])
        x, s, seen, status = eng.get_state()
        print(mode, "step", t, "x rel", rel(x, r["x"]), "sigma rel", max(rel(s[b], r["sigma"][b]) for b in range(B)), "seen", seen, r["seen"], "status", status)
