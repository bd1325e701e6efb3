"""Host-side mirror of the simulator node's per-step arithmetic (nuturtlesim/src/tube_world.cpp), batched over the C ABI, and the
device-resident closed loop it enables:  world step -> odometry -> scan -> landmarks -> associate -> update  without a host
round trip. Runs on the device; no CPU path."""
from __future__ import annotations

import numpy as np

from .nuslam import NUSLAM_DEVICE, NUSLAM_HOST, _check, lib

# nuturtlesim/config/tube_world_params.yaml, nuturtle_description/config/diff_params.yaml, nuslam/config/scan_params.yaml
TUBES = np.array([[0.5, 0.5], [-0.5, -0.5], [1.0, 1.0], [-1.0, -1.0], [-0.75, 0.75], [0.75, -0.75]], dtype=np.float64)
TUBE_RADIUS, ROBOT_RADIUS = 0.0381, 0.08
WHEEL_BASE, WHEEL_RAD = 0.16, 0.033
MIN_RANGE, MAX_RANGE = 0.05, 1.0


class TubeWorld:
    """B independent simulated robots in one tube field: one ``TubeWorld::main_loop`` iteration per ``step``
    (tube_world.cpp:512-537). ``world`` rows are {wheelBase, wheelRad, x, y, th, thL, thR, jointL, jointR}.
    ``device_resident=True`` keeps everything in torch CUDA tensors (``world``, ``ranges``, ``joints``)."""

    def __init__(self, batch, config=None, tubes=TUBES, tube_rad=TUBE_RADIUS, robot_rad=ROBOT_RADIUS, max_range=MAX_RANGE,
                 wheel_base=WHEEL_BASE, wheel_rad=WHEEL_RAD, dt=0.1, device=0, device_resident=False):
        self.batch, self.device, self.dt = int(batch), device, float(dt)
        self.tube_rad, self.robot_rad, self.max_range = float(tube_rad), float(robot_rad), float(max_range)
        w = np.zeros((self.batch, 9))
        w[:, 0], w[:, 1] = wheel_base, wheel_rad
        if config is not None:
            w[:, 2:5] = np.asarray(config, dtype=np.float64)
        tubes = np.ascontiguousarray(tubes, dtype=np.float64).reshape(-1, 2)
        self.resident = bool(device_resident)
        if self.resident:
            import torch
            dev = torch.device("cuda", device)
            self.world = torch.tensor(w, device=dev)
            self.tubes = torch.tensor(tubes, device=dev)
            self.ranges = torch.empty((self.batch, 360), dtype=torch.float32, device=dev)
            self.joints = torch.empty((2, self.batch), dtype=torch.float64, device=dev)
        else:
            self.world, self.tubes = w, tubes
            self.ranges = np.empty((self.batch, 360), dtype=np.float32)
            self.joints = np.empty((2, self.batch))

    @staticmethod
    def _p(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def step(self, cmd, noise=None, stream=None):
        """cmd [B,3] commanded twists (dth, dx, dy); noise [B,4] = (twist dth, twist dx, slip L, slip R) draws or None.
        Updates ``world`` / ``joints`` in place and returns ``ranges`` [B,360] float32."""
        if self.resident:
            if noise is not None and not noise.is_cuda:
                raise ValueError("device-resident world: pass CUDA tensors")
            mem = NUSLAM_DEVICE
        else:
            cmd = np.ascontiguousarray(np.broadcast_to(np.asarray(cmd, dtype=np.float64), (self.batch, 3)))
            noise = None if noise is None else np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), (self.batch, 4)))
            mem = NUSLAM_HOST
        _check(lib().nuslam_world_step(self._p(self.world), self._p(cmd), None if noise is None else self._p(noise), self.dt,
                                       self._p(self.tubes), int(self.tubes.shape[0]), self.tube_rad, self.robot_rad, self.max_range,
                                       self._p(self.ranges), self._p(self.joints), self.batch, mem, self.device, stream), "nuslam_world_step")
        return self.ranges

    @property
    def config(self):
        return self.world[:, 2:5]


class ClosedLoop:
    """Simulator -> odometry -> landmarks -> SLAM for B robots, every buffer resident in HBM (torch CUDA tensors):
         ranges, joints = TubeWorld.step(cmd, noise)                               tube_world.cpp:512-537
         twist = odometry.getTwist(joints); odometry(joints)                       slam.cpp:264-265
         BatchedExtendedKalman.scan_step(twist, ranges)                            landmarks.cpp:84-109 + slam.cpp:262-319
       All launches go to the EKF handle's stream, in order."""

    def __init__(self, batch, n_landmarks=12, Q=None, R=None, mode="fast", config=None, min_range=MIN_RANGE, max_range=MAX_RANGE,
                 max_markers=12, device=0, **world_kw):
        import torch
        from .nuslam import BatchedExtendedKalman
        dev = torch.device("cuda", device)
        self._torch_stream = torch.cuda.Stream(device=dev)   # ONE stream for the three stages: in-order, no events needed
        self.stream = self._torch_stream.cuda_stream
        self.world = TubeWorld(batch, config=config, device=device, device_resident=True, **world_kw)
        # the slam node starts its filter and its odometry model at the origin (slam.cpp:81-83,157)
        self.ekf = BatchedExtendedKalman(np.zeros((batch, 3)), n_landmarks=n_landmarks, Q=Q, R=R, mode=mode, device=device, stream=self.stream)
        self.min_range, self.max_range, self.m = float(min_range), float(max_range), int(max_markers)
        self.odom = torch.zeros((batch, 7), dtype=torch.float64, device=dev)
        self.odom[:, 0:2] = self.world.world[:, 0:2]
        self.twists = torch.empty((batch, 3), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)

    def step(self, cmd, noise=None):
        w = self.world
        w.step(cmd, noise, stream=self.stream)
        _check(lib().nuslam_diffdrive_step(self.odom.data_ptr(), w.joints[0].data_ptr(), w.joints[1].data_ptr(), self.twists.data_ptr(),
                                           w.batch, NUSLAM_DEVICE, w.device, self.stream), "nuslam_diffdrive_step")
        self.ekf.scan_step(self.twists, w.ranges, self.min_range, self.max_range, self.m)
