#!/usr/bin/env python
"""SURVEY.md 8(f) rows 1 + 2: the device-resident closed loop  simulator step -> odometry -> scan -> landmarks -> associate -> update
for B simulated robots (one TubeWorld + one landmarks node + one slam node each), nothing but the commanded twists crossing to the
device. Prints robot-steps/s, the per-stage split (CUDA events) and the tracking error against the simulator's ground truth."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from shermbot_navigation_b200 import tube_world  # noqa: E402
from shermbot_navigation_b200.nuslam import NUSLAM_DEVICE, _check, lib  # noqa: E402


def main(B=65536, steps=50, warmup=40, mode="fast"):
    loop = tube_world.ClosedLoop(B, n_landmarks=12, Q=0.1 * np.eye(3), R=0.001 * np.eye(2), mode=mode, max_markers=8)
    dev = torch.device("cuda")
    g = np.random.default_rng(1)
    # every robot drives its own circle (angular rate 0.1 .. 0.3 rad / step, 7 mm / step)
    cmd = torch.tensor(np.stack([g.uniform(0.1, 0.3, B), np.full(B, 0.07), np.zeros(B)], axis=1), device=dev)
    st = loop._torch_stream
    for _ in range(warmup):   # landmarks are opened here (strict-kernel first touches); the timed steps are steady state
        loop.step(cmd)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(st)
    for _ in range(steps):
        loop.step(cmd)
    t1.record(st)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    # per-stage split: the same loop with event records between the stages (one stream: the records serialise nothing)
    w = loop.world
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    for k in range(steps):
        ev[k][0].record(st)
        w.step(cmd, None, stream=loop.stream)
        ev[k][1].record(st)
        _check(lib().nuslam_diffdrive_step(loop.odom.data_ptr(), w.joints[0].data_ptr(), w.joints[1].data_ptr(), loop.twists.data_ptr(), B,
                                           NUSLAM_DEVICE, w.device, loop.stream), "nuslam_diffdrive_step")
        ev[k][2].record(st)
        loop.ekf.scan_step(loop.twists, w.ranges, loop.min_range, loop.max_range, loop.m)
        ev[k][3].record(st)
    torch.cuda.synchronize()
    split = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev]).mean(axis=0)
    x, _, seen, status = loop.ekf.get_state()
    true = w.world[:, 2:5].cpu().numpy()
    err = np.abs(x[:, 1:3] - true[:, 0:2]).max(axis=1)
    print(json.dumps({"workload": f"closed loop: {B} robots, 6 tubes, 360-beam scans, 12-landmark maps, unknown association ({mode})",
                      "ms_per_step": ms, "robot_steps_per_s": B / (ms * 1e-3),
                      "stage_ms": {"world_step": split[0], "odometry": split[1], "scan_step(detect+associate+update)": split[2]},
                      "landmarks_seen_mean": float(seen.mean()), "position_error_m": {"median": float(np.median(err)), "max": float(err.max())},
                      "bad_status": int((status != 0).sum()),
                      "algorithmic_bytes_per_robot_step": 2 * 1440 + 2 * 72 + 2 * 8 * (27 + 27 * 27) + 200}))


if __name__ == "__main__":
    main(mode=sys.argv[1] if len(sys.argv) > 1 else "fast")
