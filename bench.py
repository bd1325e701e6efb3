#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched EKF-SLAM hot path (BASELINE.json metric).

Metric: EKF predict+update steps/sec at 64K filters x 12 landmarks (fp64, known correspondence), reported
as filter-steps/s (one filter-step = 1 predict + 12 sequential updates = one iteration of
nuslam/src/slam.cpp:262-319 with 12 markers) and as a fraction of the measured HBM roofline.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU implementation, host cores

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_LANDMARKS = 12
FILTERS_PER_GPU = 65536
LEN = 3 + 2 * N_LANDMARKS
# algorithmic bytes per filter-step: read x 216 + Sigma 5832 + u 16 + z 192 + ids 48, write x 216 + Sigma 5832
BYTES_PER_FILTER_STEP = 2 * 8 * (LEN + LEN * LEN) + 16 + 20 * N_LANDMARKS
METRIC = "EKF predict+update filter-steps/s (64K filters x 12 landmarks per GPU, fp64, known correspondence)"
UNIT = "filter-steps/s"


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML every ~4 ms in a thread; nvidia-smi as fallback)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        self.max_sm = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _read_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            pw = float("nan")
        try:
            rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            try:
                rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                rs = 0
        return sm, pw, rs

    def _read_smi(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.max_sm = float(out[1])
        return float(out[0]), float(out[2]), int(out[3].strip(), 16)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self._read_nvml() if self.nvml else self._read_smi())
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        sm = [v[0] for v in self.samples]
        pw = [v[1] for v in self.samples if v[1] == v[1]]
        bits = 0
        for v in self.samples:
            bits |= v[2]
        reasons = sorted(nm for b, nm in self.REASONS.items() if bits & b)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_sm, "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": reasons, "source": "nvml" if self.nvml else "nvidia-smi"}


def make_inputs_device(torch, dev, B, steps_total, seed):
    """Synthetic config-2 inputs generated on the device (data: synthetic): robot on the reference circle
    (d_theta 0.02, dx 0.007 per step), 12 landmarks on the benign ring, exact range/bearing + N(0, 0.01^2)."""
    from shermbot_navigation_b200 import synth
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    tw1 = synth.wheel_twists(steps_total + 1, first_step_straight=True)
    poses = synth.true_trajectory(tw1)
    lm = synth.landmark_ring(N_LANDMARKS, 0.20)
    dxl = lm[None, :, 0] - poses[:, None, 1]
    dyl = lm[None, :, 1] - poses[:, None, 2]
    rng_true = torch.tensor(np.sqrt(dxl * dxl + dyl * dyl), device=dev)                      # (T, n)
    brg_true = torch.tensor(synth.wrap_pi(np.arctan2(dyl, dxl) - poses[:, None, 0]), device=dev)
    robot0 = (torch.randn((B, 3), generator=g, device=dev, dtype=torch.float64) * 0.01).cpu().numpy()
    twists = torch.tensor(tw1, device=dev)[:, None, :].expand(-1, B, -1).contiguous()           # (T, B, 3)
    ids = torch.arange(1, N_LANDMARKS + 1, device=dev, dtype=torch.int32)[None, :].expand(B, -1).contiguous()

    def z_of(t):
        noise = torch.randn((B, N_LANDMARKS, 2), generator=g, device=dev, dtype=torch.float64) * 0.01
        z = torch.empty((B, N_LANDMARKS, 2), device=dev, dtype=torch.float64)
        z[..., 0] = rng_true[t][None, :] + noise[..., 0]
        b = brg_true[t][None, :] + noise[..., 1]
        z[..., 1] = torch.atan2(torch.sin(b), torch.cos(b))
        return z

    return robot0, twists, ids, z_of


def bind_to_gpu_numa_node(index: int):
    """Run this rank's host threads on the CPUs next to its GPU (NVML's ideal affinity), so that the page-locked staging buffers of
    the end-to-end leg are allocated on that NUMA node and their DMA does not cross the socket link. Best effort; returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = len(os.sched_getaffinity(0))
        return f"nvml ideal affinity: {after} of {before} cpus"
    except Exception as e:   # containers often forbid it: not an error
        return f"unchanged ({type(e).__name__})"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from shermbot_navigation_b200 import nuslam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.filters, args.steps, args.warmup
    total_steps = W + K + 6
    robot0, twists, ids, z_of = make_inputs_device(torch, dev, B, total_steps, seed=1234 + rank)
    stream = torch.cuda.Stream(dev)          # a real (non-NULL) stream shared by torch events and the engine's launches
    torch.cuda.set_stream(stream)
    eng = nuslam.BatchedExtendedKalman(robot0, None, n_landmarks=N_LANDMARKS, mode=args.mode, device=local, stream=stream.cuda_stream)
    # state lives in torch tensors so that the final gather is a plain NCCL collective on them
    xs = torch.zeros((B, LEN), device=dev, dtype=torch.float64)
    sig = torch.zeros((B, LEN, LEN), device=dev, dtype=torch.float64)
    seen = torch.zeros(B, device=dev, dtype=torch.int32)
    status = torch.zeros(B, device=dev, dtype=torch.int32)
    x0, s0, n0, st0 = eng.get_state()
    xs.copy_(torch.from_numpy(x0))
    sig.copy_(torch.from_numpy(np.ascontiguousarray(np.transpose(s0, (0, 2, 1)))))
    torch.cuda.synchronize(dev)
    eng.bind_state(xs, sig, seen, status)
    zs = [z_of(t) for t in range(total_steps)]          # resident in HBM before the timed region
    torch.cuda.synchronize(dev)

    # step 0 touches every landmark for the first time (strict arithmetic inside the kernel); then W warm-up steps
    t = 0
    eng.step(twists[t], zs[t], ids)
    t += 1
    for _ in range(W):
        eng.step(twists[t], zs[t], ids)
        t += 1
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    ev0.record(stream)
    for _ in range(K):
        eng.step(twists[t], zs[t], ids)
        t += 1
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    bad = int((status != 0).sum().item()) + int((~torch.isfinite(xs)).any(dim=1).sum().item())

    # ---- end to end through the public API with HOST buffers (pinned): every step copies its twists / z / ids host -> device
    # and its resulting state vector device -> host inside the timed region (BatchedExtendedKalman.step_async = the C ABI's
    # nuslam_ekf_step_async: three streams, three steps in flight, so the copies of neighbouring steps overlap the kernel) ----
    Ke = max(3, min(K, args.e2e_steps))
    nbuf = 3
    h_tw = [torch.empty((B, 3), dtype=torch.float64).pin_memory() for _ in range(nbuf)]
    h_z = [torch.empty((B, N_LANDMARKS, 2), dtype=torch.float64).pin_memory() for _ in range(nbuf)]
    h_ids = ids.cpu().pin_memory()
    h_x = [torch.empty((B, LEN), dtype=torch.float64).pin_memory() for _ in range(nbuf)]
    for k in range(nbuf):
        h_tw[k].copy_(twists[t + k])
        h_z[k].copy_(zs[t + k])
    torch.cuda.synchronize(dev)
    for k in range(6):   # warm the host path (staging buffers, streams, events)
        eng.step_async(h_tw[k % nbuf].numpy(), h_z[k % nbuf].numpy(), h_ids.numpy(), h_x[k % nbuf].numpy())
    eng.wait_async()
    if world > 1:
        dist.barrier()
    checksum = 0.0
    t0 = time.perf_counter()
    for k in range(Ke):
        eng.step_async(h_tw[k % nbuf].numpy(), h_z[k % nbuf].numpy(), h_ids.numpy(), h_x[k % nbuf].numpy())
        # the host consumes the result of the step that has just left the pipeline (three calls back): robot pose of filter 0
        if k >= nbuf:
            checksum += float(h_x[k % nbuf][0, 1])
    eng.wait_async()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / float(e2e_t.item())
    h2d = B * (3 * 8 + N_LANDMARKS * 2 * 8 + N_LANDMARKS * 4)
    d2h = B * LEN * 8

    # ---- the only communication of the whole run: gather final states + error statistics (NCCL over NVLink) ----
    if world > 1:
        from shermbot_navigation_b200 import shard
        gathered = shard.gather_states(xs, world * B)
        assert gathered.shape == (world * B, LEN)
        stats = shard.allreduce_stats(torch.tensor([float(bad)], device=dev, dtype=torch.float64))
        bad = int(stats.item())

    out = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        per_launch_s = ms_max / 1e3 / K
        achieved = B * BYTES_PER_FILTER_STEP / per_launch_s / 1e9
        traffic = None
        prof = ROOT / "profiles" / "ncu_summary.json"
        if prof.exists():
            try:
                traffic = json.loads(prof.read_text()).get("ekf_step_traffic_bytes_per_launch")
            except Exception:
                pass
        value = world * B * K / (ms_max / 1e3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: batched 64K independent EKF filters x 12 landmarks, known correspondence, fp64",
                       "filters_per_gpu": B, "landmarks": N_LANDMARKS, "state_len": LEN, "mode": args.mode,
                       "batched_steps_per_s": value / (world * B) if B else None,
                       "l2": f"inputs larger than L2: filter state {B * (LEN + LEN * LEN) * 8 / 1e6:.0f} MB per GPU is streamed every step (L2 126 MB)",
                       "parallelism": f"filters sharded {world}x, no data-path collective; final NCCL all_gather of states"},
            "clocks": clocks, "gpu_launches": (2 * K if args.mode == "fast" else K), "bad_filters": bad,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "api": "BatchedExtendedKalman.step_async (nuslam_ekf_step_async): pinned host buffers, H2D + fused step + D2H of the state vector per step, 3 steps in flight", "host_affinity": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_filter_step": BYTES_PER_FILTER_STEP,
                         "kernel": "k_ekf_fast_step<12> (+ k_ekf_strict_list over the first-touch work list, empty in steady state)" if args.mode == "fast" else "k_ekf_strict<kOpStep>",
                         "launch_us": per_launch_s * 1e6, "launches_per_step": 2 if args.mode == "fast" else 1},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return out


def cpu_reference_run(n_filters, n_steps, nthreads):
    """Time the reference's own CPU implementation (oracle/_ref when built, else the C restatement) on a bounded
    sample of the same workload: n_filters filters x n_steps fused steps after the first-touch step."""
    import oracle
    from shermbot_navigation_b200 import synth
    orc = oracle.best()
    sc = synth.ekf_scenario(n_filters, n_steps + 1, n=N_LANDMARKS, seed=4321)
    first = orc.ekf_run(N_LANDMARKS, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][:1], sc["z"][:1], sc["ids"][:1], nthreads=nthreads)
    t0 = time.perf_counter()
    orc.ekf_run(N_LANDMARKS, sc["robot0"], sc["map0"], sc["Q"], sc["R"], sc["twists"][1:], sc["z"][1:], sc["ids"][1:],
                init=(first["x"], first["sigma"], first["seen"]), nthreads=nthreads)
    dt = time.perf_counter() - t0
    return n_filters * n_steps / dt, dt, ("reference" if orc.kind.startswith("ref") else "port")


def cpu_baseline(target_seconds=12.0):
    cores = os.cpu_count() or 1
    # calibrate on a tiny run, then size the sample for ~target_seconds of wall time on all cores
    rate, _, kind = cpu_reference_run(4 * cores, 10, cores)
    n_steps = 50
    n_filters = int(max(cores, min(FILTERS_PER_GPU, rate * target_seconds / n_steps)))
    n_filters = max(cores, (n_filters // cores) * cores)
    rate, dt, kind = cpu_reference_run(n_filters, n_steps, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n_filters} filters x {n_steps} fused steps (12 landmarks, known correspondence) on {cores} threads, {dt:.1f} s; "
                      "naive-loop Armadillo shim, -O2 -ffp-contract=off"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # each "step" = one batched fused step over a bounded sample of the 64K-filter workload
    rate, _, kind = cpu_reference_run(4 * cores, 5, cores)
    per_step_budget = min(2.0, 150.0 / max(1, K + W))
    n_filters = int(max(cores, min(FILTERS_PER_GPU, rate * per_step_budget)))
    n_filters = max(cores, (n_filters // cores) * cores)
    import oracle
    from shermbot_navigation_b200 import synth
    orc = oracle.best()
    sc = synth.ekf_scenario(n_filters, 2, n=N_LANDMARKS, seed=4321)
    # a persistent batch of reference filters: the state stays inside the oracle library between steps, so a timed step is the
    # reference's arithmetic on all host threads and nothing else (no per-step copies on the Python side)
    run = orc.ekf_stepper(N_LANDMARKS, sc["robot0"], sc["map0"], sc["Q"], sc["R"], nthreads=cores)
    tw = [np.ascontiguousarray(sc["twists"][t]) for t in range(2)]
    zz = [np.ascontiguousarray(sc["z"][t]) for t in range(2)]
    ii = [np.ascontiguousarray(sc["ids"][t], dtype=np.int32) for t in range(2)]
    run.step(tw[0], zz[0], ii[0])   # the first-touch step
    times = []
    for k in range(W + K):
        t0 = time.perf_counter()
        run.step(tw[1], zz[1], ii[1])
        dt = time.perf_counter() - t0
        if k >= W:
            times.append(dt)
    total = float(np.sum(times))
    value = n_filters * K / total
    sample = (f"{n_filters} of {FILTERS_PER_GPU} filters per step x {K} steps on {cores} host threads "
              f"({'oracle/_ref: unmodified reference sources' if kind == 'reference' else 'oracle C restatement'}, naive-loop Armadillo shim, -O2)")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: batched 64K independent EKF filters x 12 landmarks, known correspondence, fp64 "
                               "(reference CPU implementation timed on a bounded sample)", "filters_per_step": n_filters, "landmarks": N_LANDMARKS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--filters", type=int, default=FILTERS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="ekf", choices=["ekf", "scan", "assoc", "large", "closed_loop"],
                    help="ekf (default): the headline line of BASELINE.json configs[1]; the others run the per-config benches under tools/ "
                         "(config 3 scans, config 4 shard with on-device association, config 5 large map, the device-resident closed loop) "
                         "on one GPU and print their own JSON line")
    args = ap.parse_args()
    if args.config != "ekf":
        import runpy
        tool = {"scan": "bench_scan.py", "assoc": "bench_assoc.py", "large": "bench_large.py", "closed_loop": "bench_closed_loop.py"}[args.config]
        sys.argv = [str(ROOT / "tools" / tool)]
        runpy.run_path(str(ROOT / "tools" / tool), run_name="__main__")
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
