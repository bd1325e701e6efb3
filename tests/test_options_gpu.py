"""The opt-in departures from the reference (NUSLAM_OPT_*: wrapped innovation, Joseph form, pre-motion Jacobian, finite landmark
prior) against a dense numpy fp64 restatement of the same formulas (tests/sane_reference.py) -- the reference has none
of them, so there is no compiled oracle here -- plus the properties they exist for. Tolerance 1e-9 relative."""
import numpy as np
import pytest

from sane_reference import JOSEPH, PRE_MOTION, WRAP, SaneEkf
from shermbot_navigation_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(a).max(), np.abs(b).max(), 1e-300)


def run_both(nuslam, sc, B, T, options, prior, mode="strict", known=True):
    n = sc["n"]
    eng = nuslam.BatchedExtendedKalman(sc["robot0"][:B], n_landmarks=n, Q=sc["Q"], R=sc["R"], mode=mode, options=options, landmark_prior=prior)
    refs = [SaneEkf(n, sc["robot0"][b], sc["Q"], sc["R"], prior=prior if prior is not None else 2147483647.0, options=options) for b in range(B)]
    ids_ok = True
    for t in range(T):
        got = eng.step(sc["twists"][t, :B], sc["z"][t, :B], sc["ids"][t, :B] if known else None, return_ids=True)
        for b, f in enumerate(refs):
            want = f.step(sc["twists"][t, b], sc["z"][t, b], sc["ids"][t, b] if known else None)
            ids_ok &= list(got[b]) == want
    x, s, seen, status = eng.get_state()
    return x, s, seen, status, refs, ids_ok


def test_reference_behaviour_is_the_default(cuda_lib):
    """options = 0, prior = INT_MAX: the numpy restatement is the reference's algorithm; after the first touches both agree loosely
    (the INT_MAX cancellation amplifies rounding) -- a sanity check of the restatement itself."""
    sc = synth.ekf_scenario(4, 6, n=6, seed=2)
    x, s, seen, status, refs, ids_ok = run_both(cuda_lib, sc, 4, 6, 0, None)
    assert ids_ok and not status.any()
    assert max(rel(x[b], refs[b].x) for b in range(4)) < 1e-3


@pytest.mark.parametrize("options", [WRAP, JOSEPH, PRE_MOTION, WRAP | JOSEPH | PRE_MOTION])
@pytest.mark.parametrize("geometry", ["benign", "adversarial"])
def test_options_match_numpy(cuda_lib, options, geometry):
    B, T, n = 6, 25, 12
    sc = synth.ekf_scenario(B, T, n=n, seed=31, geometry=geometry, shuffle_order=True)
    x, s, seen, status, refs, ids_ok = run_both(cuda_lib, sc, B, T, options, 100.0)
    assert ids_ok and not status.any()
    ex = max(rel(x[b], refs[b].x) for b in range(B))
    es = max(rel(s[b], refs[b].S) for b in range(B))
    print(f"[options {options} {geometry}] x rel {ex:.2e}, Sigma rel {es:.2e}")
    assert ex < TOL and es < TOL
    if options & JOSEPH:
        for b in range(B):
            assert np.array_equal(s[b], s[b].T), "Joseph form returns an exactly symmetric Sigma"
            assert np.linalg.eigvalsh(s[b]).min() > -1e-12


def test_options_with_association_and_fast_mode_request(cuda_lib):
    """Unknown association with wrapped innovations; a FAST-mode handle with options set runs the oracle-order kernels (same results)."""
    B, T, n = 5, 12, 12
    sc = synth.ekf_scenario(B, T, n=n, seed=7, geometry="adversarial", shuffle_order=True)
    res = {}
    for mode in ("strict", "fast"):
        x, s, seen, status, refs, ids_ok = run_both(cuda_lib, sc, B, T, WRAP | JOSEPH, 50.0, mode=mode, known=False)
        assert ids_ok and not status.any()
        assert max(rel(x[b], refs[b].x) for b in range(B)) < TOL
        res[mode] = (x, s)
    assert np.array_equal(res["strict"][0], res["fast"][0]) and np.array_equal(res["strict"][1], res["fast"][1])


def test_finite_prior_removes_the_first_touch_asymmetry(cuda_lib):
    """What the finite prior is for: with INT_MAX the reference's first update leaves Sigma asymmetric at ~1e-6;
    with a finite prior the same arithmetic stays symmetric to rounding."""
    sc = synth.ekf_scenario(8, 3, n=12, seed=5)
    out = {}
    for prior in (None, 100.0):
        eng = cuda_lib.BatchedExtendedKalman(sc["robot0"], n_landmarks=12, Q=sc["Q"], R=sc["R"], mode="strict", landmark_prior=prior)
        for t in range(3):
            eng.step(sc["twists"][t], sc["z"][t], sc["ids"][t])
        _, s, _, _ = eng.get_state()
        out[prior] = max(np.abs(s[b] - s[b].T).max() / np.abs(s[b]).max() for b in range(8))
    print(f"[prior] relative asymmetry of Sigma after 3 steps: INT_MAX {out[None]:.2e}, prior 100 {out[100.0]:.2e}")
    assert out[None] > 1e-9 and out[100.0] < 1e-11
