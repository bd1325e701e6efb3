"""Host-side mirror of the slice of ``rigid2d`` the EKF caller protocol uses (rigid2d/include/rigid2d/diff_drive.hpp:13-103),
batched over the C ABI: ``DiffDrive.getTwist`` + ``DiffDrive.__call__`` (what nuslam/src/slam.cpp:264-265 does with the wheel
angles of a joint-state message) and ``convertTwist``. Runs on the device; no CPU path."""
from __future__ import annotations

import numpy as np

from .nuslam import NUSLAM_HOST, _check, lib


def integrateTwist(twists, device=0):
    """rigid2d::integrateTwist (rigid2d.cpp:294-328), batched: twists [N,3] (dth, dx, dy) -> [N,4] = (cos, sin, x, y) of the
    Transform2D the twist reaches in unit time."""
    tw = np.ascontiguousarray(np.atleast_2d(np.asarray(twists, dtype=np.float64)))
    out = np.empty((tw.shape[0], 4))
    _check(lib().nuslam_integrate_twist(tw.ctypes.data, out.ctypes.data, tw.shape[0], NUSLAM_HOST, device, None), "nuslam_integrate_twist")
    return out


class DiffDrive:
    """B differential-drive robots. ``config`` rows are (x, y, th); wheel angles start at (thL, thR) = 0 unless given."""

    def __init__(self, wheel_base, wheel_rad, config=None, batch=None, thL=0.0, thR=0.0, device=0):
        if config is None:
            config = np.zeros((batch or 1, 3))
        config = np.atleast_2d(np.asarray(config, dtype=np.float64))
        B = config.shape[0]
        self.device = device
        self.state = np.zeros((B, 7))
        self.state[:, 0] = wheel_base
        self.state[:, 1] = wheel_rad
        self.state[:, 2:5] = config
        self.state[:, 5] = thL
        self.state[:, 6] = thR

    def step(self, thL_new, thR_new):
        """getTwist(thL, thR) then operator()(thL, thR) (diff_drive.cpp:80-146): returns the body twists [B,3] = (dth, dx, 0)."""
        B = self.state.shape[0]
        l = np.ascontiguousarray(np.broadcast_to(np.asarray(thL_new, dtype=np.float64), (B,)))
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(thR_new, dtype=np.float64), (B,)))
        tw = np.empty((B, 3))
        _check(lib().nuslam_diffdrive_step(self.state.ctypes.data, l.ctypes.data, r.ctypes.data, tw.ctypes.data, B, NUSLAM_HOST, self.device, None),
               "nuslam_diffdrive_step")
        return tw

    def convertTwist(self, twists):
        """DiffDrive::convertTwist (diff_drive.cpp:66-78): twists [B,3] -> wheel velocities [B,2] (uL, uR)."""
        tw = np.ascontiguousarray(np.atleast_2d(np.asarray(twists, dtype=np.float64)))
        u = np.empty((tw.shape[0], 2))
        _check(lib().nuslam_diffdrive_convert_twist(float(self.state[0, 0]), float(self.state[0, 1]), tw.ctypes.data, u.ctypes.data, tw.shape[0],
                                                    NUSLAM_HOST, self.device, None), "nuslam_diffdrive_convert_twist")
        return u

    @property
    def config(self):
        return self.state[:, 2:5]
