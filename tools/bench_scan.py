#!/usr/bin/env python
"""BASELINE.json config 3: S synthetic 360-beam scans resident in HBM -> clustering + classification + circle fit.
Prints scans/s and achieved GB/s against the algorithmic bytes (1440 in + 720 cluster ids + 8 + 32 per published circle)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from shermbot_navigation_b200 import circle_fit, synth  # noqa: E402


def main(S=1_000_000, distinct=16384, steps=5, warmup=2):
    sd = synth.scan_scenario(distinct, seed=101, noise_sigma=0.001)
    reps = (S + distinct - 1) // distinct
    r = torch.tensor(np.tile(sd["ranges"], (reps, 1))[:S], device="cuda")
    for _ in range(warmup):
        out = circle_fit.scan_detect(r, sd["min_range"], sd["max_range"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        out = circle_fit.scan_detect(r, sd["min_range"], sd["max_range"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ncirc = float(out["n_circles"].clamp(min=0).double().mean())
    bytes_per_scan = 1440 + 720 + 8 + 32 * ncirc
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    gbs = S * bytes_per_scan / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": f"config 3: {S} scans x 360 beams (noise 1 mm), clustering + classification + Jacobi circle fit", "ms": ms,
                      "scans_per_s": S / (ms * 1e-3), "mean_circles_per_scan": ncirc, "mean_clusters_per_scan": float(out["n_clusters"].double().mean()),
                      "algorithmic_bytes_per_scan": bytes_per_scan, "achieved_gbs": gbs, "frac_of_hbm": gbs / peak}))


if __name__ == "__main__":
    main()
