"""Plain numpy fp64 restatement of the EKF step WITH the opt-in departures from the reference (NUSLAM_OPT_*).
The reference itself has none of them, so the compiled oracle cannot check them; with options = 0 and prior = INT_MAX this class is
the reference's algorithm (slam_library.cpp:65-282, slam.cpp:262-319) in dense matrices. Test infrastructure only."""
import math

import numpy as np

WRAP, JOSEPH, PRE_MOTION = 1, 2, 4


def normalize_angle(a):
    return math.atan2(math.sin(a), math.cos(a))


class SaneEkf:
    def __init__(self, n, robot, Q, R, prior=2147483647.0, options=0, amin=0.01, amax=60.0):
        self.n, self.len = n, 3 + 2 * n
        self.x = np.zeros(self.len)
        self.x[:3] = robot
        self.S = np.zeros((self.len, self.len))
        for i in range(3, self.len):
            self.S[i, i] = prior
        self.Q, self.R = np.asarray(Q, float), np.asarray(R, float)
        self.opt, self.seen, self.amin, self.amax = options, 0, amin, amax

    def predict(self, dth, dx):
        th = self.x[0]
        if dth == 0.0:
            dq = (0.0, dx * math.cos(th), dx * math.sin(th))
        else:
            q = dx / dth
            dq = (dth, -q * math.sin(th) + q * math.sin(th + dth), q * math.cos(th) - q * math.cos(th + dth))
        self.x[0] += dq[0]
        self.x[1] += dq[1]
        self.x[2] += dq[2]
        thJ = th if (self.opt & PRE_MOTION) else self.x[0]
        A = np.eye(self.len)
        if dth == 0.0:
            A[1, 0], A[2, 0] = -dx * math.sin(thJ), dx * math.cos(thJ)
        else:
            q = dx / dth
            A[1, 0] = -q * math.cos(thJ) + q * math.cos(thJ + dth)
            A[2, 0] = -q * math.sin(thJ) + q * math.sin(thJ + dth)
        self.S = A @ self.S @ A.T
        self.S[:3, :3] += self.Q

    def model(self, j):
        c = 3 + 2 * (j - 1)
        dx, dy = self.x[c] - self.x[1], self.x[c + 1] - self.x[2]
        d = dx * dx + dy * dy
        sq = math.sqrt(d)
        zhat = np.array([sq, normalize_angle(normalize_angle(math.atan2(dy, dx)) - self.x[0])])
        H = np.zeros((2, self.len))
        H[1, 0] = -1.0
        H[0, 1], H[1, 1] = -dx / sq, dy / d
        H[0, 2], H[1, 2] = -dy / sq, -dx / d
        H[0, c], H[1, c] = dx / sq, -dy / d
        H[0, c + 1], H[1, c + 1] = dy / sq, dx / d
        return zhat, H

    def innovation(self, z, zhat):
        dz = np.asarray(z, float) - zhat
        if self.opt & WRAP:
            dz[1] = normalize_angle(dz[1])
        return dz

    def associate(self, z):
        if self.seen == 0:
            self.seen = 1
            return 1
        assert 3 + 2 * self.seen < self.len, "map full"
        for k in range(1, self.seen + 1):
            zhat, H = self.model(k)
            psi = H @ self.S @ H.T + self.R
            dz = self.innovation(z, zhat)
            d = float(dz @ np.linalg.inv(psi) @ dz)
            if d < self.amin:
                return k
            if self.amin < d < self.amax:
                return -1
        self.seen += 1
        return self.seen

    def init_landmark(self, z, j):
        c = 3 + 2 * (j - 1)
        self.x[c] = self.x[1] + z[0] * math.cos(z[1] + self.x[0])
        self.x[c + 1] = self.x[2] + z[0] * math.sin(z[1] + self.x[0])

    def update(self, z, j):
        zhat, H = self.model(j)
        K = self.S @ H.T @ np.linalg.inv(H @ self.S @ H.T + self.R)
        self.x = self.x + K @ self.innovation(z, zhat)
        self.x[0] = normalize_angle(self.x[0])
        M = np.eye(self.len) - K @ H
        if self.opt & JOSEPH:
            S = M @ self.S @ M.T + K @ self.R @ K.T
            self.S = 0.5 * (S + S.T)
        else:
            self.S = M @ self.S

    def step(self, tw, zs, ids=None):
        """One iteration of slam.cpp:262-319; ids None = associateLandmark. Returns the ids used."""
        snapshot = self.seen
        self.predict(tw[0], tw[1])
        out = []
        for i, z in enumerate(zs):
            if ids is None:
                j = self.associate(z)
            else:
                j = int(ids[i])
                if j <= 0:
                    out.append(0)
                    continue
                self.seen = max(self.seen, j)
            out.append(j)
            if j > snapshot:
                self.init_landmark(z, j)
            elif j < 0:
                continue
            self.update(z, j)
        return out
