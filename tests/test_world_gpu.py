"""GPU parity tests of the two rows SURVEY.md 8(f) ranks next: the batched simulator step (nuturtlesim/src/tube_world.cpp)
and the fused scan -> landmarks -> associate -> update step (landmarks.cpp:84-109 into slam.cpp:262-319), both through the
C ABI against the oracle.

Tolerances: integer outputs (marker counts, association ids) bit-exact; fp64 state 1e-9 relative under teacher forcing
(both sides start every step from the same state and the SAME measurements: the EKF's first touch of a landmark amplifies
any input difference); float32 ranges identical up to the CUDA-vs-glibc atan2/sincos ulp (reported,
bounded at 1 float ulp on < 0.1 % of the beams)."""
import numpy as np
import pytest

from shermbot_navigation_b200 import tube_world

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel_max(a, b):
    scale = max(np.abs(a).max(), np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def random_world(B, seed):
    g = np.random.default_rng(seed)
    cfg = np.stack([g.uniform(-1.2, 1.2, B), g.uniform(-1.2, 1.2, B), g.uniform(-3.1, 3.1, B)], axis=1)
    cmd = np.stack([g.uniform(-0.5, 0.5, B), g.uniform(0.0, 0.2, B), np.zeros(B)], axis=1)
    cmd[::7, 0] = 0.0   # straight-line robots: the dth == 0 branch of integrateTwist
    return cfg, cmd, g


def test_world_step_matches_oracle(cuda_lib, orc):
    B, T = 1024, 6
    cfg, cmd, g = random_world(B, 11)
    # a few robots start inside a tube's collision radius (check_collision) and exactly on a tube's axis (horizontal ray: dy = 0)
    cfg[:6, 0:2] = tube_world.TUBES + np.array([0.05, -0.04])
    cfg[6, 0:3] = [0.0, 0.5, 0.0]
    w = tube_world.TubeWorld(B, config=cfg)
    wo = w.world.copy()
    exact = beams = 0
    worst_ulp = 0
    for t in range(T):
        noise = np.stack([g.normal(0, 0.01, B), g.normal(0, 0.01, B), g.normal(0.95, 0.05, B), g.normal(0.95, 0.05, B)], axis=1)
        if t == 0:
            noise[:] = 0.0
        r = w.step(cmd, noise).copy()
        ro = orc.world_step(wo, cmd, noise, w.dt, tube_world.TUBES, w.tube_rad, w.robot_rad, w.max_range)
        assert rel_max(w.world, wo) < 1e-13
        ulp = np.abs(r.view(np.int32).astype(np.int64) - ro.view(np.int32).astype(np.int64))
        exact += int((ulp == 0).sum())
        beams += ulp.size
        worst_ulp = max(worst_ulp, int(ulp.max()))
        wo[:] = w.world   # teacher-forced: the next step starts from identical poses (int(rad2deg(th)) is a cliff)
    print(f"[world] beams identical: {exact}/{beams}, worst difference {worst_ulp} float ulp; hits per scan {np.mean((r < 1.5).sum(axis=1)):.1f}")
    assert worst_ulp <= 1 and exact >= 0.999 * beams
    assert (r < 1.5).any()


def oracle_scan_step(orc, n, Q, R, state, twists, z, mv):
    """One slam.cpp:262-319 iteration per filter with unknown association, filter b using its first mv[b] measurements."""
    x, s, seen = state
    B, m = z.shape[0], z.shape[1]
    xo, so, no = x.copy(), s.copy(), seen.copy()
    ids = np.zeros((B, m), dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    for k in np.unique(mv):
        sel = np.nonzero(mv == k)[0]
        zz = z[sel][None, :, :max(k, 1)] if k > 0 else np.zeros((1, sel.size, 0, 2))
        r = orc.ekf_run(n, np.zeros((sel.size, 3)), np.zeros((sel.size, 2 * n)), Q, R, twists[sel][None], zz, None,
                        init=(x[sel], s[sel], seen[sel]))
        xo[sel], so[sel], no[sel], status[sel] = r["x"], r["sigma"], r["seen"], r["status"]
        if k > 0:
            ids[sel, :k] = r["ids_out"][0]
    return xo, so, no, ids, status


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_scan_step_matches_oracle(cuda_lib, orc, mode):
    """scan -> markers -> z -> associate -> [init] -> update, fused; world scans of a robot circling between the tubes."""
    B, T, n, m = 96, 14, 12, 8
    g = np.random.default_rng(5)
    cfg = np.stack([g.uniform(-0.3, 0.3, B), g.uniform(-0.3, 0.3, B), g.uniform(-3.1, 3.1, B)], axis=1)
    world = np.zeros((B, 9))
    world[:, 0], world[:, 1], world[:, 2:5] = tube_world.WHEEL_BASE, tube_world.WHEEL_RAD, cfg
    Q, R = 0.1 * np.eye(3), 0.001 * np.eye(2)
    eng = cuda_lib.BatchedExtendedKalman(cfg, n_landmarks=n, Q=Q, R=R, mode=mode)
    cmd = np.tile(np.array([0.2, 0.07, 0.0]), (B, 1))
    odom = world[:, :7].copy()
    new_landmarks = used = 0
    for t in range(T):
        ranges = orc.world_step(world, cmd, None, 0.1, tube_world.TUBES, tube_world.TUBE_RADIUS, tube_world.ROBOT_RADIUS, 1.0)
        if t == 3:
            ranges[::5] = 2.0          # scans without any marker: predict only
            ranges[1, :] = 0.5         # one all-equal scan (clusterPoints' undefined wrap case or a single huge cluster)
        tw = np.zeros((B, 3))
        # odometry twists (slam.cpp:264-265) from the encoder readings, through the oracle
        for b in range(B):
            st, twb = orc.diffdrive_step(odom[b], world[b, 7], world[b, 8])
            odom[b], tw[b] = st, twb
        x0, s0, seen0, _ = eng.get_state()
        nm, z, ids = eng.scan_step(tw, ranges, tube_world.MIN_RANGE, tube_world.MAX_RANGE, m, return_all=True)
        xg, sg, seeng, stg = eng.get_state()
        # stage 1: the markers (landmarks.cpp:84-109) and their polar form (slam.cpp:282-286)
        det = orc.scan_detect_batch(ranges, tube_world.MIN_RANGE, tube_world.MAX_RANGE, kmax=m)
        assert np.array_equal(nm, det["n_circles"])
        mv = np.clip(nm, 0, m)
        for b in range(B):
            for i in range(mv[b]):
                zo = orc.cartesian2polar(det["circles"][b, i, 0], det["circles"][b, i, 1])
                assert np.abs(z[b, i] - zo).max() <= 1e-9 * max(1.0, np.abs(zo).max())
            assert not z[b, mv[b]:].any()
        # stage 2: the EKF iteration, teacher-forced on the state and fed the SAME measurements
        xo, so, no, ido, sto = oracle_scan_step(orc, n, Q, R, (x0, s0, seen0), tw, z, mv)
        assert np.array_equal(ids, ido), f"association ids differ at step {t}"
        assert np.array_equal(seeng, no) and np.array_equal(stg != 0, sto != 0)
        assert rel_max(xg, xo) < TOL
        for b in range(B):
            assert rel_max(sg[b], so[b]) < TOL
        new_landmarks += int((no - seen0).sum())
        used += int(mv.sum())
    print(f"[scan_step {mode}] measurements used {used}, landmarks opened {new_landmarks}, seen per filter {seeng.mean():.2f}")
    assert new_landmarks > B and used > 4 * B


def test_closed_loop_device_resident(cuda_lib):
    """world -> odometry -> detect -> associate -> update with every buffer in HBM: the filter tracks the simulated robot."""
    import torch
    B, T = 512, 120
    loop = tube_world.ClosedLoop(B, n_landmarks=12, Q=0.1 * np.eye(3), R=0.001 * np.eye(2), mode="fast", max_markers=8)
    cmd = torch.tensor(np.tile(np.array([0.2, 0.07, 0.0]), (B, 1)), device="cuda")
    for t in range(T):
        loop.step(cmd)
    loop.ekf.synchronize()
    x, _, seen, status = loop.ekf.get_state()
    true = loop.world.world[:, 2:5].cpu().numpy()
    err = np.abs(x[:, 1:3] - true[:, 0:2]).max()
    print(f"[closed loop] {T} steps, landmarks seen {seen.mean():.2f}, max position error {err:.3e} m, status bits {np.bitwise_or.reduce(status)}")
    assert seen.min() >= 1 and err < 0.05


def test_map_to_odom_matches_oracle(cuda_lib, orc):
    """EKFSlam::broadcast_map2odom_tf (slam.cpp:175-210), the step after the path: T_mo = T_mb * T_ob.inv()."""
    B, n = 257, 12
    g = np.random.default_rng(8)
    est = np.stack([g.uniform(-3.14, 3.14, B), g.uniform(-2, 2, B), g.uniform(-2, 2, B)], axis=1)
    eng = cuda_lib.BatchedExtendedKalman(est, n_landmarks=n, mode="fast")
    odom = np.zeros((B, 7))
    odom[:, 0], odom[:, 1] = tube_world.WHEEL_BASE, tube_world.WHEEL_RAD
    odom[:, 2:5] = np.stack([g.uniform(-2, 2, B), g.uniform(-2, 2, B), g.uniform(-6.5, 6.5, B)], axis=1)   # odometry theta is never normalised
    odom[0, 2:5] = est[0, [1, 2, 0]]   # identical frames: the identity transform
    got = eng.map_to_odom(odom)
    want = np.stack([orc.map_to_odom(odom[b, 2:5], est[b]) for b in range(B)])
    assert np.abs(got - want).max() < 1e-14
    assert np.abs(got[0]).max() < 1e-15


def test_replay_matches_closed_loop(cuda_lib, tmp_path):
    """The ROS-free replay of the slam node's loop (joint positions + scans in, estimates + map->odom out) reproduces the
    device-resident closed loop that produced the log, bit for bit; also through its command line."""
    import subprocess
    import sys
    import torch
    from shermbot_navigation_b200 import replay
    B, T = 48, 30
    Q, R = 0.1 * np.eye(3), 0.001 * np.eye(2)
    loop = tube_world.ClosedLoop(B, n_landmarks=12, Q=Q, R=R, mode="fast", max_markers=8)
    cmd = torch.tensor(np.tile(np.array([0.15, 0.06, 0.0]), (B, 1)), device="cuda")
    joints, ranges = np.empty((T, 2, B)), np.empty((T, B, 360), np.float32)
    for t in range(T):
        loop.step(cmd)
        loop.ekf.synchronize()
        joints[t] = loop.world.joints.cpu().numpy()
        ranges[t] = loop.world.ranges.cpu().numpy()
    xl, sl, seenl, stl = loop.ekf.get_state()
    res = replay.replay(joints, ranges, n_landmarks=12, Q=Q, R=R, mode="fast", max_markers=8)
    x, s, seen, st = res["state"]
    assert np.array_equal(x, xl) and np.array_equal(s, sl) and np.array_equal(seen, seenl) and np.array_equal(st, stl)
    assert np.array_equal(res["pose"][-1], x[:, :3])
    # map -> odom composed with odom -> body gives the estimate back: T_mo * T_ob = T_mb
    tx, ty, yaw = res["map_to_odom"][-1].T
    ox, oy, oth = res["odom"][-1].T
    bx = tx + np.cos(yaw) * ox - np.sin(yaw) * oy
    by = ty + np.sin(yaw) * ox + np.cos(yaw) * oy
    assert np.abs(bx - x[:, 1]).max() < 1e-12 and np.abs(by - x[:, 2]).max() < 1e-12
    np.savez(tmp_path / "log.npz", joints=joints, ranges=ranges)
    r = subprocess.run([sys.executable, "-m", "shermbot_navigation_b200.replay", str(tmp_path / "log.npz"), str(tmp_path / "out.npz")],
                       capture_output=True, text=True, cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert r.returncode == 0, r.stderr
    # the command line uses the node's defaults (Q = 0.1 I, R = 0.001 I, 12 markers): same filter, same log
    out = np.load(tmp_path / "out.npz")
    assert out["pose"].shape == (T, B, 3) and np.isfinite(out["x"]).all()


def test_closed_loop_free_running_vs_oracle(cuda_lib, orc):
    """L2-style check of the whole fused stack: the device-resident closed loop against the SAME loop built from oracle calls
    (world step -> DiffDrive odometry -> scan_detect -> cartesian2polar -> associate / initialize / update), free running from
    scratch for 80 steps. First touches amplify rounding, so the bound is loose; association sequences
    and tracking quality must agree."""
    import torch
    B, T, n, m = 24, 80, 12, 8
    g = np.random.default_rng(17)
    cmd = np.stack([g.uniform(0.1, 0.3, B), np.full(B, 0.07), np.zeros(B)], axis=1)
    Q, R = 0.1 * np.eye(3), 0.001 * np.eye(2)
    loop = tube_world.ClosedLoop(B, n_landmarks=n, Q=Q, R=R, mode="fast", max_markers=m)
    cmd_d = torch.tensor(cmd, device="cuda")
    # oracle side
    world = np.zeros((B, 9))
    world[:, 0], world[:, 1] = tube_world.WHEEL_BASE, tube_world.WHEEL_RAD
    odom = world[:, :7].copy()
    filt = [orc.ekf(n, np.zeros(3), np.zeros(2 * n), Q, R) for _ in range(B)]
    same_ids = np.ones(B, dtype=bool)
    for t in range(T):
        loop.step(cmd_d)
        ranges = orc.world_step(world, cmd, None, 0.1, tube_world.TUBES, tube_world.TUBE_RADIUS, tube_world.ROBOT_RADIUS, 1.0)
        det = orc.scan_detect_batch(ranges, tube_world.MIN_RANGE, tube_world.MAX_RANGE, kmax=m)
        for b, f in enumerate(filt):
            odom[b], tw = orc.diffdrive_step(odom[b], world[b, 7], world[b, 8])
            _, _, snapshot = f.get()
            f.predict(tw[0], tw[1], tw[2])
            for i in range(min(max(det["n_circles"][b], 0), m)):
                z = orc.cartesian2polar(det["circles"][b, i, 0], det["circles"][b, i, 1])
                j = f.associate(z)
                if j > snapshot:
                    f.init_landmark(z, j)
                elif j < 0:
                    continue
                f.update(z, j)
    loop.ekf.synchronize()
    x, s, seen, status = loop.ekf.get_state()
    xo = np.stack([f.get()[0] for f in filt])
    seeno = np.array([f.get()[2] for f in filt])
    true = world[:, 2:5]
    assert rel_max(loop.world.world.cpu().numpy(), world) < 1e-12          # the simulators agree
    err_gpu = np.abs(x[:, 1:3] - true[:, 0:2]).max(axis=1)
    err_orc = np.abs(xo[:, 1:3] - true[:, 0:2]).max(axis=1)
    diff = np.abs(x[:, :3] - xo[:, :3]).max(axis=1)
    agree = (seen == seeno) & (diff < 1e-3)
    print(f"[closed loop vs oracle] {T} steps: robots agreeing {agree.sum()}/{B}, worst pose difference among them {diff[agree].max():.2e}; "
          f"tracking error median gpu {np.median(err_gpu):.2e} / oracle {np.median(err_orc):.2e}, max gpu {err_gpu.max():.2e} / oracle {err_orc.max():.2e}")
    assert agree.sum() >= 0.9 * B and not status.any()
    assert abs(np.median(err_gpu) - np.median(err_orc)) < 1e-3
