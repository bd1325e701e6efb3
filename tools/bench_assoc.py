#!/usr/bin/env python
"""BASELINE.json config 4: Monte-Carlo EKF-SLAM with UNKNOWN data association (associateLandmark on the device, Mahalanobis gating)
for `total` filters x 12 landmarks sharded over the ranks of the job (total / world filters per GPU, no data-path collective),
steady state (all landmarks seen). `run()` returns whole-job filter-steps/s, the roofline of the shard's kernel, the error
statistics reduced on the device and summed over the ranks by one NCCL all-reduce, and the reference's CPU path beside it."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

N = 12
LEN = 3 + 2 * N
BYTES_PER_FILTER_STEP = 2 * 8 * (LEN + LEN * LEN) + 16 + 16 * N + 4 + 4 * N   # SURVEY.md 8d with on-device association: 12 356 B


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def cpu_assoc(seconds):
    """The reference's own EKF with associateLandmark (slam.cpp:262-319) on all host threads, steady state, bounded sample."""
    import oracle
    from shermbot_navigation_b200 import synth
    cores = os.cpu_count() or 1
    best = None
    for kind in ("ref_blas", "ref", "port"):
        if not oracle.available(kind):
            continue
        orc = oracle.load(kind)
        sc = synth.ekf_scenario(4 * cores, 8, n=N, seed=99, shuffle_order=True)
        head = orc.ekf_run(N, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][:2], sc["z"][:2], None, nthreads=cores)
        t0 = time.perf_counter()
        orc.ekf_run(N, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][2:], sc["z"][2:], None,
                    init=(head["x"], head["sigma"], head["seen"]), nthreads=cores)
        rate = 4 * cores * 6 / (time.perf_counter() - t0)
        steps = 20
        nf = int(max(cores, min(65536, rate * seconds / steps)) // cores * cores)
        sc = synth.ekf_scenario(nf, steps + 2, n=N, seed=98, shuffle_order=True)
        head = orc.ekf_run(N, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][:2], sc["z"][:2], None, nthreads=cores)
        t0 = time.perf_counter()
        orc.ekf_run(N, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][2:], sc["z"][2:], None,
                    init=(head["x"], head["sigma"], head["seen"]), nthreads=cores)
        dt = time.perf_counter() - t0
        r = {"value": nf * steps / dt, "unit": "filter-steps/s", "cores": cores, "kind": "reference" if kind.startswith("ref") else "port",
             "variant": orc.flavour, "sample": f"{nf} filters x {steps} fused steps with associateLandmark on {cores} threads, {dt:.1f} s"}
        if best is None or r["value"] > best["value"]:
            best = r
        if kind == "ref":
            break   # the restatement only stands in when the compiled reference is absent
    return best


def run(total=1 << 20, steps=10, warmup=3, mode="fast", cpu_seconds=6.0, world=1, rank=0, dist=None):
    import torch
    from shermbot_navigation_b200 import nuslam, synth
    B = total // world
    D = 256
    T = warmup + steps + 1
    sc = synth.ekf_scenario(D, T, n=N, seed=9 + rank, shuffle_order=True)
    dev = torch.device("cuda", torch.cuda.current_device())
    rep_np = lambda a: np.ascontiguousarray(np.tile(a, (B // D,) + (1,) * (a.ndim - 1)))
    rep = lambda a: torch.tensor(a, device=dev).repeat((B // D,) + (1,) * (a.ndim - 1))
    stream = torch.cuda.Stream(dev)
    prev_stream = torch.cuda.current_stream(dev)
    torch.cuda.set_stream(stream)
    try:
        eng = nuslam.BatchedExtendedKalman(rep_np(sc["robot0"]), rep_np(sc["map0"]), sc["Q"], sc["R"], mode=mode, device=dev.index, stream=stream.cuda_stream)
        tw = [rep(sc["twists"][t]) for t in range(T)]
        z = [rep(sc["z"][t]) for t in range(T)]
        truth = [rep(sc["ids"][t].astype(np.int32)) for t in range(T)]
        # step 0 builds the map with associateLandmark itself: ids are assigned in detection order; truth -> assigned
        ids0 = eng.step(tw[0], z[0], None, return_ids=True)
        perm = torch.zeros((B, N + 1), dtype=torch.int32, device=dev)
        perm.scatter_(1, truth[0].long(), ids0)
        for t in range(1, 1 + warmup):
            eng.step(tw[t], z[t], None)
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        got = None
        for t in range(1 + warmup, 1 + warmup + steps):
            got = eng.step(tw[t], z[t], None, return_ids=(t == warmup + steps))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms_t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        # K6: error statistics of the shard, reduced on the device; the ranks add them up with ONE all-reduce (the run's only collective)
        t_last = warmup + steps
        want = torch.gather(perm, 1, truth[t_last].long())
        pose = torch.tensor(np.ascontiguousarray(np.broadcast_to(sc["poses"][t_last], (B, 3))), device=dev)
        # landmark j of the estimate is the j-th landmark in detection order of step 0: compare landmark positions through the same map
        stats = eng.error_stats(truth_pose=pose, truth_map=None, ids_got=got, ids_want=want)
        if dist is not None:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(dev)
        ms = float(ms_t.item())
        st = dict(zip(nuslam.BatchedExtendedKalman.STATS, [float(v) for v in stats.cpu().numpy()]))
        eng.close()
    finally:
        torch.cuda.set_stream(prev_stream)
    nf = max(st["filters"], 1.0)
    peak = hbm_peak()
    per_gpu = B / (ms * 1e-3)
    gbs = per_gpu * BYTES_PER_FILTER_STEP / 1e9
    res = {"workload": f"BASELINE.json configs[3]: {total}-filter Monte-Carlo EKF-SLAM x {N} landmarks, unknown data association (Mahalanobis gating), "
                       f"sharded {world}x ({B} filters per GPU), steady state",
           "value": world * per_gpu, "unit": "filter-steps/s", "scaling": "strong", "n_gpus": world, "filters_per_gpu": B, "ms_per_step": ms, "mode": mode,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "algorithmic_bytes_per_filter_step": BYTES_PER_FILTER_STEP, "kernel": "k_ekf_fast_step<12, BULK, ASSOC> (per GPU)"},
           "stats": {"reduced": "k_error_stats on every rank + one NCCL all-reduce(sum)" if dist is not None else "k_error_stats",
                     "filters": st["filters"], "rmse_position_m": (st["sq_position_error"] / nf) ** 0.5, "rmse_heading_rad": (st["sq_heading_error"] / nf) ** 0.5,
                     "mean_nees_3dof": st["nees"] / nf, "bad_status": st["bad_status"],
                     "ids_differing_from_truth_last_step": st["id_mismatches"], "decisions_last_step": world * B * N,
                     "note": "the reference returns -1 (ambiguous) for 0.01 < d < 60, so ids can differ from the truth by design; parity with the oracle's ids is tested in tests/"}}
    if cpu_seconds > 0 and rank == 0:
        res["cpu_baseline"] = cpu_assoc(cpu_seconds)
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    import torch
    torch.cuda.set_device(0)
    print(json.dumps(run(total=int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20)))
