#!/usr/bin/env python
"""BASELINE.json config 3: S synthetic 360-beam scans resident in HBM -> clustering + classification + circle fit.
`run()` returns scans/s, achieved GB/s against the algorithmic bytes (1440 in + 720 cluster ids + 8 + 32 per published circle) and,
beside it, the reference's CPU path (oracle/_ref: clusterPoints -> classifyCluster -> circleFit, all host threads) on a bounded sample."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def cpu_scans(ranges, min_range, max_range, seconds):
    """The reference's own scan path on all host threads, on a sample sized for ~`seconds` of wall time."""
    import oracle
    orc = oracle.best()
    cores = os.cpu_count() or 1
    probe = ranges[:256 * cores] if len(ranges) >= 256 * cores else ranges
    t0 = time.perf_counter()
    orc.scan_detect_batch(probe, min_range, max_range, nthreads=0)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(max(cores, min(len(ranges), rate * seconds)))
    t0 = time.perf_counter()
    orc.scan_detect_batch(ranges[:n], min_range, max_range, nthreads=0)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "scans/s", "cores": cores, "kind": "reference" if orc.kind.startswith("ref") else "port",
            "sample": f"{n} of the same scans on {cores} threads, {dt:.1f} s; unmodified reference sources, Armadillo shim (Jacobi svd / eig_sym)"}


def run(S=1_000_000, distinct=16384, steps=5, warmup=2, cpu_seconds=6.0, fit="moment"):
    import torch
    from shermbot_navigation_b200 import circle_fit, synth
    sd = synth.scan_scenario(distinct, seed=101, noise_sigma=0.001)
    reps = (S + distinct - 1) // distinct
    r = torch.tensor(np.tile(sd["ranges"], (reps, 1))[:S], device="cuda")
    prev = circle_fit.set_fit(fit)
    try:
        for _ in range(warmup):
            out = circle_fit.scan_detect(r, sd["min_range"], sd["max_range"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            out = circle_fit.scan_detect(r, sd["min_range"], sd["max_range"])
        e1.record()
        torch.cuda.synchronize()
        fallbacks = circle_fit.last_fallbacks() if fit == "moment" else None
    finally:
        circle_fit.set_fit(prev)
    ms = e0.elapsed_time(e1) / steps
    ncirc = float(out["n_circles"].clamp(min=0).double().mean())
    bytes_per_scan = 1440 + 720 + 8 + 32 * ncirc
    peak = hbm_peak()
    gbs = S * bytes_per_scan / (ms * 1e-3) / 1e9
    res = {"workload": f"BASELINE.json configs[2]: {S} synthetic 360-beam scans (1 mm range noise), clustering + circle classification + algebraic circle fit",
           "value": S / (ms * 1e-3), "unit": "scans/s", "ms_per_pass": ms, "fit": fit, "scans_rerun_in_oracle_order": fallbacks,
           "mean_circles_per_scan": ncirc, "mean_clusters_per_scan": float(out["n_clusters"].double().mean()),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "algorithmic_bytes_per_scan": bytes_per_scan,
                        "kernel": "k_scan_moment (+ k_scan_detect<true> over the re-run list)" if fit == "moment" else "k_scan_detect / k_scan_fit_* / k_scan_publish"}}
    if cpu_seconds > 0:
        res["cpu_baseline"] = cpu_scans(sd["ranges"], sd["min_range"], sd["max_range"], cpu_seconds)
    del r, out
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    print(json.dumps(run(fit=sys.argv[1] if len(sys.argv) > 1 else "moment")))
