import sys, time, faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, '/root/repo')
import numpy as np
import oracle
from shermbot_navigation_b200 import circle_fit, synth
orc = oracle.best()
which = sys.argv[1]
n = int(sys.argv[2])
sd = synth.scan_scenario(max(n, 1), seed=41, noise_sigma=0.0)
r = sd["ranges"] if which != "edge" else synth.edge_scans()
print("mode", which, "scans", len(r), flush=True)
circle_fit.set_fit("moment")
t0 = time.time()
got = circle_fit.scan_detect(r, 0.05, 1.0)
print("gpu done", time.time() - t0, "fallbacks", circle_fit.last_fallbacks(), flush=True)
want = orc.scan_detect_batch(r, 0.05, 1.0, nthreads=0)
print("cob equal", np.array_equal(got["cluster_of_beam"], want["cluster_of_beam"]), "ncl", np.array_equal(got["n_clusters"], want["n_clusters"]),
      "nci", np.array_equal(got["n_circles"], want["n_circles"]), flush=True)
bad = np.nonzero(got["n_circles"] != want["n_circles"])[0]
print("scans with differing n_circles:", bad[:10], got["n_circles"][bad[:10]], want["n_circles"][bad[:10]])
k = max(1, int(want["n_circles"].clip(0).max()))
a, b = got["circles"][:, :k, :3], want["circles"][:, :k, :3]
ok = want["n_circles"] == got["n_circles"]
err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
mask = (np.arange(k)[None, :] < want["n_circles"].clip(0)[:, None]) & ok[:, None]
print("worst circle rel err", np.nanmax(np.where(mask[..., None], err, 0.0)))
