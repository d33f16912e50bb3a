"""CPU: the NDT refinement stage of the restated matcher (matchScanCustom.m:32-50).

MathWorks' matchScans is not in the reference tree (PARITY UNPINNED), so these
tests pin what can be pinned without it: the call contract of the wrapper
(accept rule, validity gate, covariance of the grid stage), the internal
consistency of the restatement (analytic gradient / Hessian against finite
differences, continuity where a point changes NDT blocks, monotone ascent) and
that the stage is off by default (the golden vectors were made without it)."""
import numpy as np
import pytest

import oracle as O

GUESS = np.array([0.6, 0.15, 0.1])


def scan(G, i):
    return O.Scan(G["intel_ranges"][int(i)], G["intel_angles"])


@pytest.fixture()
def seeded(golden):
    m = O.Map()
    for p, si in zip(golden["upd_seed_poses"], golden["upd_seed_scan_idx"]):
        m.update(p, scan(golden, si))
    return m, scan(golden, 4)


@pytest.fixture()
def refine_on():
    old = O.set_refine(True)
    yield
    O.set_refine(old)


def grid_optimum(m, s):
    """Correction (cells, cells, rad) found by the grid stage from GUESS."""
    b = m.match(GUESS, s, 0.7, 0.7)["best"]
    return np.array([b[0], b[1], b[2] * O.rot_step()])


def test_stage_is_off_by_default(seeded):
    m, s = seeded
    o = m.match(GUESS, s, 0.5, 0.5)
    assert o["ndt_evals"] == 0 and not o["ndt_accepted"]
    assert o["score"] == int(o["score"])


def test_gradient_and_hessian_match_finite_differences(seeded):
    m, s = seeded
    rng = np.random.default_rng(3)
    p0 = grid_optimum(m, s)
    for _ in range(4):
        p = p0 + rng.normal(0, [0.4, 0.4, 0.003])
        t = O.ndt_terms(m, GUESS, s, p)
        assert t["S"] > 1.0
        h = np.array([1e-5, 1e-5, 1e-7])
        g_fd = np.empty(3)
        H_fd = np.empty((3, 3))
        for a in range(3):
            e = np.zeros(3)
            e[a] = h[a]
            tp, tm = O.ndt_terms(m, GUESS, s, p + e), O.ndt_terms(m, GUESS, s, p - e)
            g_fd[a] = (tp["S"] - tm["S"]) / (2 * h[a])
            H_fd[a] = (tp["grad"] - tm["grad"]) / (2 * h[a])
        assert np.allclose(t["grad"], g_fd, rtol=1e-5, atol=1e-5 * np.abs(t["grad"]).max())
        scale = np.sqrt(np.outer(np.abs(np.diag(t["hess"])), np.abs(np.diag(t["hess"]))))
        assert np.all(np.abs(t["hess"] - H_fd) < 1e-4 * scale + 1e-6)
        # the curvature model is symmetric positive semi-definite
        assert np.linalg.eigvalsh(t["model"]).min() > -1e-9 * np.abs(t["model"]).max()


def test_score_is_continuous_in_the_correction(seeded):
    """Points change NDT blocks at half-integer lattice coordinates; the C1 window makes
    the score continuous there (a truncated Gaussian would jump by up to exp(-0.5))."""
    m, s = seeded
    p0 = grid_optimum(m, s)
    xs = np.linspace(-1.0, 1.0, 401)
    S = np.array([O.ndt_terms(m, GUESS, s, [p0[0] + x, p0[1] + 0.3, p0[2] + 0.002])["S"] for x in xs])
    assert S.max() > 5.0
    assert np.abs(np.diff(S)).max() < 0.02 * S.max()
    # and the step sizes are those of a smooth curve: second differences stay small
    assert np.abs(np.diff(S, 2)).max() < 0.01 * S.max()


def test_refine_contract(seeded, refine_on):
    m, s = seeded
    rng = np.random.default_rng(15)
    n_acc = 0
    for _ in range(12):
        g = GUESS + rng.normal(0, [0.15, 0.15, 0.08])
        rx = ry = float(rng.uniform(0.25, 0.7))
        O.set_refine(False)
        grid = m.match(g, s, rx, ry)
        O.set_refine(True)
        o = m.match(g, s, rx, ry)
        assert o["best"] == grid["best"] and o["valid"] == grid["valid"]
        assert np.array_equal(o["cov"], grid["cov"], equal_nan=True)           # matchScanCustom.m:22 -- never the NDT stage's
        if not grid["valid"]:
            assert o["ndt_evals"] == 0 and o["score"] == 0.0                   # :26-28 returns before matchScans
            continue
        assert 1 <= o["ndt_evals"] <= 500                                     # :36 MaxIterations
        if o["ndt_accepted"]:
            n_acc += 1
            d = o["pose"] - g
            assert abs(d[0]) < rx and abs(d[1]) < ry and abs(d[2]) < np.pi / 6 and np.any(d != 0)   # :38 isValidPose
            assert 2 * o["score"] > grid["score"]                             # :39
            # the refined pose is a local maximum reached by ascent from the grid pose
            step = O.rot_step()
            p0 = np.array([grid["best"][0], grid["best"][1], grid["best"][2] * step])
            p1 = np.array([d[0] / 0.05, d[1] / 0.05, d[2]])
            t0, t1 = O.ndt_terms(m, g, s, p0), O.ndt_terms(m, g, s, p1)
            assert t1["S"] >= t0["S"] and np.isclose(t1["S"], o["score"], rtol=1e-9)
            assert np.abs(np.linalg.solve(t1["model"] + 1e-9 * np.eye(3), t1["grad"]))[:2].max() < 0.05
        else:
            assert np.array_equal(o["pose"], grid["pose"]) and o["score"] == grid["score"]      # :42-45
    assert n_acc >= 4


def test_filter_runs_with_the_stage(golden, refine_on):
    """The whole per-scan update with the NDT stage on: finite poses, most matches valid."""
    N = 6
    f = O.Filter(N, 180, 30)
    rng = np.random.default_rng(2)
    f.set_scan(golden["intel_ranges"][0], golden["intel_angles"])
    f.integrate()
    f.integrate()
    for si in range(1, 6):
        f.motion(1, np.array([0.02, 0.0, 0.01, 0.0]), 1.0, np.array([0.01, 0.01, 0.01, 0.01]))
        f.set_scan(golden["intel_ranges"][si], golden["intel_angles"])
        f.map_update(rng.standard_normal((N, 30, 3)))
    assert np.isfinite(f.pose).all()


def test_known_answers_of_the_restatement():
    """tests/golden/ndt_oracle_kat.npz (made by tests/golden/make_ndt_regression.py) holds outputs of
    OUR restatement, not of MathWorks matchScans: a regression pin, bit for bit."""
    import os
    import sys

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_ndt_regression as K

    want = np.load(os.path.join(here, "ndt_oracle_kat.npz"))
    got = K.run()
    assert want["accepted"].sum() >= 3
    for k in ("valid", "evals", "accepted", "best"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["score"], want["score"])
