"""ROS-free replay of the slam node's main loop: what `nuslam/src/slam.cpp` does with the messages it receives --
wheel joint positions (sensor_msgs/JointState) and laser scans (sensor_msgs/LaserScan) -- for B robots at once, from arrays instead
of topics:

    twist = odom_model.getTwist(joints); odom_model(joints)            slam.cpp:264-265  (nuslam_diffdrive_step)
    markers = Landmarks::main_loop(scan)                               landmarks.cpp:84-109
    predict / associate / initialize / update                          slam.cpp:269-319  (nuslam_ekf_scan_step, fused with the line above)
    T_map_odom = T_map_body * T_odom_body^-1                           slam.cpp:175-210  (nuslam_ekf_map_to_odom)

Every stage runs on the device through the C ABI; there is no CPU path.

    python -m shermbot_navigation_b200.replay log.npz out.npz      # log: joints [T,2,B] f64, ranges [T,B,360] f32
"""
from __future__ import annotations

import sys

import numpy as np

from .nuslam import NUSLAM_HOST, BatchedExtendedKalman, _check, lib
from .tube_world import MAX_RANGE, MIN_RANGE, WHEEL_BASE, WHEEL_RAD


def replay(joints, ranges, n_landmarks=12, Q=None, R=None, mode="fast", max_markers=12, wheel_base=WHEEL_BASE, wheel_rad=WHEEL_RAD,
           min_range=MIN_RANGE, max_range=MAX_RANGE, device=0, options=0, landmark_prior=None):
    """joints [T,2,B] (left, right wheel angle per step), ranges [T,B,360] float32. Returns a dict of per-step outputs:
    ``pose`` [T,B,3] (theta, x, y estimate), ``odom`` [T,B,3] (x, y, theta of the odometry model), ``map_to_odom`` [T,B,3]
    (tx, ty, yaw), ``n_markers`` [T,B], ``ids`` [T,B,max_markers], and the final ``state`` (x, Sigma, seen, status)."""
    joints = np.ascontiguousarray(joints, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    T, _, B = joints.shape
    ekf = BatchedExtendedKalman(np.zeros((B, 3)), n_landmarks=n_landmarks, Q=Q, R=R, mode=mode, device=device, options=options,
                                landmark_prior=landmark_prior)   # the node starts at the origin (slam.cpp:81-83,157)
    odom = np.zeros((B, 7))
    odom[:, 0], odom[:, 1] = wheel_base, wheel_rad
    out = dict(pose=np.empty((T, B, 3)), odom=np.empty((T, B, 3)), map_to_odom=np.empty((T, B, 3)), n_markers=np.empty((T, B), np.int32),
               ids=np.empty((T, B, max_markers), np.int32))
    tw = np.empty((B, 3))
    for t in range(T):
        _check(lib().nuslam_diffdrive_step(odom.ctypes.data, joints[t, 0].ctypes.data, joints[t, 1].ctypes.data, tw.ctypes.data, B, NUSLAM_HOST,
                                           device, None), "nuslam_diffdrive_step")
        nm, _, ids = ekf.scan_step(tw, ranges[t], min_range, max_range, max_markers, return_all=True)
        x = ekf.getStateVector()
        out["pose"][t] = x[:, :3]
        out["odom"][t] = odom[:, 2:5]
        out["map_to_odom"][t] = ekf.map_to_odom(odom)
        out["n_markers"][t] = nm
        out["ids"][t] = ids
    out["state"] = ekf.get_state()
    return out


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    log = np.load(argv[1])
    res = replay(log["joints"], log["ranges"])
    x, sigma, seen, status = res.pop("state")
    np.savez_compressed(argv[2], x=x, sigma=sigma, seen=seen, status=status, **res)
    print(f"replayed {log['joints'].shape[0]} steps for {log['joints'].shape[2]} robots -> {argv[2]}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
