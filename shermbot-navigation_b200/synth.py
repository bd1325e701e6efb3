"""Synthetic inputs for the EKF-SLAM and scan-detection paths (host side, numpy).

The reference has no reproducible data source: its simulator seeds ``std::mt19937`` from
``random_device`` and steps on wall-clock time (nuturtlesim/src/tube_world.cpp:57-58,522-523).
This module generates the deterministic stand-ins described in SURVEY.md 8(d), from the
reference's own constants:

* TurtleBot3 burger kinematics: wheel_radius 0.033, wheel_base 0.16
  (nuturtle_description/config/diff_params.yaml:2-3)
* Q = 0.1 I3, R = 0.001 I2 (nuslam/config/slam_params.yaml:2-3)
* lidar: 360 beams, range window [0.05, 1.0], fill value max+1 (nuturtlesim/config/scan_params.yaml:1-2,
  nuturtlesim/src/tube_world.cpp:416), tube radius 0.0381 and the six tube positions
  (nuturtlesim/config/tube_world_params.yaml:2-9)

Everything is generated once on the host from a counter-keyed Philox stream and handed
identically to the oracle and to the GPU engine.
"""
from __future__ import annotations

import numpy as np

WHEEL_RADIUS = 0.033
WHEEL_BASE = 0.16
Q_DEFAULT = np.diag([0.1, 0.1, 0.1])
R_DEFAULT = np.diag([0.001, 0.001])
TUBE_RADIUS = 0.0381
TUBES = np.array([[0.5, 0.5], [-0.5, -0.5], [1.0, 1.0], [-1.0, -1.0], [-0.75, 0.75], [0.75, -0.75]])
MIN_RANGE = 0.05
MAX_RANGE = 1.0
SCAN_FILL = MAX_RANGE + 1.0
MEAS_SIGMA = 0.01  # nuturtle_description/urdf/turtlebot3_burger.gazebo.xacro:125


def _rng(seed, *key):
    k = np.array([seed & 0xFFFFFFFFFFFFFFFF, hash(tuple(int(v) for v in key)) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
    return np.random.Generator(np.random.Philox(key=k))


def wrap_pi(a):
    """(-pi, pi] wrap used only to build plausible measurements (not the reference's normalize_angle)."""
    return np.arctan2(np.sin(a), np.cos(a))


def landmark_ring(n=12, radius=0.20, centre=(0.0, 0.35), phase=0.1):
    """benign set: ring of radius 0.20 about the centre of the driven circle (bearings stay in [0.9, 2.2] rad);
    adversarial set: radius 0.60 (bearings cross +-pi)."""
    ang = phase + 2.0 * np.pi * np.arange(n) / n
    return np.stack([centre[0] + radius * np.cos(ang), centre[1] + radius * np.sin(ang)], axis=1)


def wheel_twists(T, dth=0.02, dx=0.007, first_step_straight=True):
    """Constant body twist realised as wheel increments through convertTwist and recovered with getTwist
    (rigid2d/src/diff_drive.cpp:66-110), so the control fed to predict is what slam.cpp:264 would see."""
    d = WHEEL_BASE / 2
    r = WHEEL_RADIUS
    tw = np.zeros((T, 3))
    thL = thR = 0.0
    for t in range(T):
        om = 0.0 if (first_step_straight and t == 0) else dth
        uL = (-(d / r) * om) + (dx / r)
        uR = ((d / r) * om) + (dx / r)
        nL, nR = thL + uL, thR + uR
        dUL, dUR = nL - thL, nR - thR
        tw[t, 0] = (r / WHEEL_BASE) * (dUR - dUL)
        tw[t, 1] = (r / 2) * (dUL + dUR)
        thL, thR = nL, nR
    return tw


def true_trajectory(twists):
    """Exact unicycle integration of the twists from the origin; returns poses AFTER each step (T,3) = (th,x,y)."""
    T = twists.shape[0]
    pose = np.zeros(3)
    out = np.zeros((T, 3))
    for t in range(T):
        dth, dx = twists[t, 0], twists[t, 1]
        th = pose[0]
        if dth == 0.0:
            pose = pose + np.array([0.0, dx * np.cos(th), dx * np.sin(th)])
        else:
            rr = dx / dth
            pose = pose + np.array([dth, -rr * np.sin(th) + rr * np.sin(th + dth), rr * np.cos(th) - rr * np.cos(th + dth)])
        out[t] = pose
    return out


def ekf_scenario(B, T, n=12, seed=12345, geometry="benign", start_sigma=0.01, meas_sigma=MEAS_SIGMA,
                 first_step_straight=True, shuffle_order=False, dropout=0.0):
    """Config-2 style workload: B filters, T steps, all n landmarks measured every step.

    Returns dict with twists (T,B,3), z (T,B,n,2) polar, ids (T,B,n) int32 (1-based, 0 = no measurement),
    robot0 (B,3), map0 (B,2n) zeros, landmarks (n,2), Q, R.
    """
    radius = 0.20 if geometry == "benign" else 0.60
    lm = landmark_ring(n, radius)
    tw1 = wheel_twists(T, first_step_straight=first_step_straight)
    poses = true_trajectory(tw1)
    g = _rng(seed, 1, B, T, n)
    robot0 = g.normal(0.0, start_sigma, size=(B, 3)) if start_sigma > 0 else np.zeros((B, 3))
    twists = np.broadcast_to(tw1[:, None, :], (T, B, 3)).copy()
    dxl = lm[None, :, 0] - poses[:, None, 1]
    dyl = lm[None, :, 1] - poses[:, None, 2]
    rng_true = np.sqrt(dxl * dxl + dyl * dyl)                      # (T,n)
    brg_true = wrap_pi(np.arctan2(dyl, dxl) - poses[:, None, 0])   # (T,n)
    noise = g.normal(0.0, meas_sigma, size=(T, B, n, 2)) if meas_sigma > 0 else np.zeros((T, B, n, 2))
    z = np.empty((T, B, n, 2))
    z[..., 0] = rng_true[:, None, :] + noise[..., 0]
    z[..., 1] = wrap_pi(brg_true[:, None, :] + noise[..., 1])
    ids = np.broadcast_to(np.arange(1, n + 1, dtype=np.int32)[None, None, :], (T, B, n)).copy()
    if shuffle_order:
        for t in range(T):
            for b in range(B):
                p = g.permutation(n)
                z[t, b] = z[t, b, p]
                ids[t, b] = ids[t, b, p]
    if dropout > 0:
        drop = g.random(size=(T, B, n)) < dropout
        ids[drop] = 0
    return dict(twists=twists, z=z, ids=ids, robot0=robot0, map0=np.zeros((B, 2 * n)), landmarks=lm,
                Q=Q_DEFAULT.copy(), R=R_DEFAULT.copy(), poses=poses, n=n)


# ----------------------------------------------------------------------------------------- scans

def simulate_scans(poses, tubes, tube_radius=TUBE_RADIUS, walls=None, noise_sigma=0.0, seed=0,
                   max_range=MAX_RANGE, fill=SCAN_FILL):
    """Exact ray/circle (and ray/segment) intersection lidar, 360 one-degree beams, stored as float32.

    poses (S,3) = (th,x,y); tubes: (K,2) shared or (S,K,2) per scan (NaN rows = absent); walls: optional
    (S,W,4) segments (x0,y0,x1,y1) or None. Beam i points at heading + i degrees. A hit beyond
    max_range keeps its value only if <= max_range, otherwise the fill value max_range + 1
    (tube_world.cpp:416,462-464).
    """
    poses = np.asarray(poses, dtype=np.float64)
    S = poses.shape[0]
    tubes = np.asarray(tubes, dtype=np.float64)
    if tubes.ndim == 2:
        tubes = np.broadcast_to(tubes[None], (S,) + tubes.shape)
    ang = poses[:, 0:1] + np.deg2rad(np.arange(360))[None, :]          # (S,360)
    ux, uy = np.cos(ang), np.sin(ang)
    best = np.full((S, 360), np.inf)
    for k in range(tubes.shape[1]):
        cx = tubes[:, k, 0:1] - poses[:, 1:2]
        cy = tubes[:, k, 1:2] - poses[:, 2:3]
        proj = cx * ux + cy * uy
        disc = tube_radius ** 2 - (cx * cx + cy * cy) + proj * proj
        with np.errstate(invalid="ignore"):
            t = proj - np.sqrt(disc)
        ok = (disc >= 0) & (t > 0)
        best = np.where(ok & (t < best), t, best)
    if walls is not None:
        walls = np.asarray(walls, dtype=np.float64)
        for w in range(walls.shape[1]):
            x0 = walls[:, w, 0:1] - poses[:, 1:2]
            y0 = walls[:, w, 1:2] - poses[:, 2:3]
            ex = walls[:, w, 2:3] - walls[:, w, 0:1]
            ey = walls[:, w, 3:4] - walls[:, w, 1:2]
            den = ux * ey - uy * ex
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (x0 * ey - y0 * ex) / den
                s = (x0 * uy - y0 * ux) / den
            ok = (np.abs(den) > 1e-12) & (t > 0) & (s >= 0) & (s <= 1)
            best = np.where(ok & (t < best), t, best)
    if noise_sigma > 0:
        g = _rng(seed, 2, S)
        best = best + g.normal(0.0, noise_sigma, size=best.shape)
    out = np.where(np.isfinite(best) & (best <= max_range), best, fill)
    return out.astype(np.float32)


def scan_scenario(S, seed=777, noise_sigma=0.0, wall_frac=0.10, wrap_frac=0.05, extra_tubes=6):
    """Config-3 style workload: S scans, robot pose uniform in [-1,1]^2 x (-pi,pi], the six reference
    tubes plus 0..extra_tubes random ones, wall_frac of the scans with a wall segment (non-circle cluster),
    wrap_frac with a tube dead ahead so that its cluster spans beam 359 -> 0."""
    g = _rng(seed, 3, S)
    poses = np.stack([g.uniform(-np.pi, np.pi, S), g.uniform(-1, 1, S), g.uniform(-1, 1, S)], axis=1)
    K = TUBES.shape[0] + extra_tubes + 1
    tubes = np.full((S, K, 2), np.nan)
    tubes[:, :TUBES.shape[0]] = TUBES[None]
    n_extra = g.integers(0, extra_tubes + 1, S)
    ext = g.uniform(-1.5, 1.5, size=(S, extra_tubes, 2))
    for k in range(extra_tubes):
        use = n_extra > k
        tubes[use, TUBES.shape[0] + k] = ext[use, k]
    wrap = g.random(S) < wrap_frac
    dist = g.uniform(0.25, 0.8, S)
    ahead = np.stack([poses[:, 1] + dist * np.cos(poses[:, 0]), poses[:, 2] + dist * np.sin(poses[:, 0])], axis=1)
    tubes[wrap, K - 1] = ahead[wrap]
    # robots inside a tube would see nonsense: push those tubes away
    d = np.sqrt((tubes[..., 0] - poses[:, None, 1]) ** 2 + (tubes[..., 1] - poses[:, None, 2]) ** 2)
    tubes[d < TUBE_RADIUS + 0.06] = np.nan
    tubes = np.where(np.isnan(tubes), 1e6, tubes)
    walls = np.full((S, 1, 4), 1e6)
    has_wall = g.random(S) < wall_frac
    wd = g.uniform(0.3, 0.9, S)
    wa = g.uniform(-np.pi, np.pi, S)
    half = g.uniform(0.1, 0.5, S)
    cxw = poses[:, 1] + wd * np.cos(wa)
    cyw = poses[:, 2] + wd * np.sin(wa)
    walls[has_wall, 0, 0] = (cxw - half * np.sin(wa))[has_wall]
    walls[has_wall, 0, 1] = (cyw + half * np.cos(wa))[has_wall]
    walls[has_wall, 0, 2] = (cxw + half * np.sin(wa))[has_wall]
    walls[has_wall, 0, 3] = (cyw - half * np.cos(wa))[has_wall]
    ranges = simulate_scans(poses, tubes, walls=walls, noise_sigma=noise_sigma, seed=seed)
    return dict(ranges=ranges, poses=poses, min_range=MIN_RANGE, max_range=MAX_RANGE)


def edge_scans(n_random=200, seed=3):
    """Hand-made scans for the corner cases of clusterPoints (circle_fit_library.cpp:136-206): empty scan, one open cluster
    with the wrap rule and no closed cluster (undefined behaviour in the reference), wrap onto a cluster that does not
    contain beam 0, clusters of sizes 1..5 (erase-loop skip), NaN ranges (count as in range), every beam closing a
    cluster, wrap with a dropped open tail, plus random sparse scans."""
    scans = []
    scans.append(np.full(360, 2.0, np.float32))
    scans.append(np.full(360, 0.5, np.float32))
    a = np.full(360, 2.0, np.float32)
    a[357:360] = 0.5
    a[0:3] = 0.5
    a[100:110] = 0.6
    a[110] = 0.9
    scans.append(a)
    b = np.full(360, 2.0, np.float32)
    for k in range(0, 350, 7):
        b[k:k + (k // 7) % 5 + 1] = 0.3 + 0.001 * k
    scans.append(b)
    c = np.full(360, 2.0, np.float32)
    c[10:20] = np.nan
    c[20] = 0.4
    scans.append(c)
    scans.append((0.1 + 0.05 * (np.arange(360) % 17)).astype(np.float32))
    e = np.full(360, 2.0, np.float32)
    e[0:5] = 0.5
    e[355:360] = 0.52
    scans.append(e)
    rng = np.random.default_rng(seed)
    for _ in range(n_random):
        scans.append(np.where(rng.random(360) < 0.6, rng.uniform(0.04, 1.1, 360), 2.0).astype(np.float32))
    return np.stack(scans)
