"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
reference-generated golden vectors.  Integer/byte/index results are compared
bit-exact; floating-point results with the tolerance written at each assert.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

DIM = 800
# float outputs: GPU vs oracle differ only through cos/sin/exp (CUDA libm vs glibc,
# <= 2 ulp each) and through summing log-odds as exact integers instead of float64
POSE_ATOL = 1e-11
REL_TOL = 1e-10


@pytest.fixture(scope="module")
def PS():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200.particles import ParticleSet

    return ParticleSet


def scan_of(G, i):
    return G["intel_ranges"][int(i)], G["intel_angles"]


def oscan(G, i):
    return O.Scan(*scan_of(G, i))


def tenths(a):
    return np.rint(np.asarray(a) * 10.0).astype(np.int32)


def assert_map_equal(ps, particle, omap, what=""):
    """GPU tiles of one particle == oracle tiles: same set of reference tiles,
    every cell equal in integer tenths (reference float64 = tenths/10 +- 1.1e-14)."""
    ot = omap.tiles()
    gt = ps.list_tiles(particle)
    assert sorted(gt) == sorted(ot.keys()), "%s tile sets differ: %s vs %s" % (what, gt, list(ot.keys()))
    for (cx, cy), ref in ot.items():
        got = ps.export_tile(particle, cx, cy)
        assert got is not None
        if not np.array_equal(tenths(got), tenths(ref)):
            bad = np.argwhere(tenths(got) != tenths(ref))
            raise AssertionError("%s tile (%d,%d): %d cells differ, first %s got %s want %s" % (
                what, cx, cy, len(bad), bad[0], got[tuple(bad[0])], ref[tuple(bad[0])]))
        assert np.max(np.abs(got - ref)) <= 1.2e-14 + 1e-15


def dense_tiles(G, prefix):
    out = {}
    for n, (cx, cy) in enumerate(G[prefix + "_centres"]):
        a = np.zeros(DIM * DIM)
        a[G["%s_t%d_idx" % (prefix, n)]] = G["%s_t%d_val" % (prefix, n)]
        out[(int(cx), int(cy))] = a.reshape(DIM, DIM)
    return out


# ------------------------------------------------------------------ motion --

@pytest.mark.parametrize("name,family,par", [
    ("abs", 0, (0, 0, 0, 0)),
    ("velraw", 1, (0.002, 0.05, 0.01 * np.pi / 180, 0.05)),
    ("velaces", 1, (0.02, 0.01, 0.2 * np.pi / 180, 0.02)),
    ("uni", 2, (0, 0, 0, 0)),
])
def test_motion_families(PS, golden, name, family, par):
    ps = PS(5, 180, pool_subtiles=64)
    ps.poses = golden["mot_pose0"]
    ps.covs = golden["mot_cov0"]
    dt = float(golden["mot_%s_dt_ticks" % name]) / 1e4
    ps.motion(family, golden["mot_%s_u" % name], dt, par)
    pose, cov = ps.poses, ps.covs
    op, oc = O.motion(family, golden["mot_%s_u" % name], dt, par, golden["mot_pose0"], golden["mot_cov0"])
    for i in range(5):
        assert np.array_equal(pose[i], op)                           # same sin / cos on both sides (rb_math.h)
        assert np.array_equal(cov[i], oc)
        assert np.allclose(pose[i], golden["mot_%s_pose" % name], rtol=0, atol=1e-13)
        assert np.allclose(cov[i], golden["mot_%s_cov" % name], rtol=1e-12, atol=1e-18)


# -------------------------------------------------------- map integration --

def test_integration_golden_trajectory(PS, golden):
    """hybridmap.py:95-145 across tile borders, negative coordinates, lazy tile
    allocation: GPU == oracle == Python reference."""
    ps = PS(3, 180, pool_subtiles=600)
    m = O.Map()
    for p, si in zip(golden["integ_poses"], golden["integ_scan_idx"]):
        ps.poses = p
        ps.set_scan(*scan_of(golden, si))
        ps.integrate()
        m.update(p, oscan(golden, si))
    ps.synchronize()
    for i in range(3):
        assert_map_equal(ps, i, m, "integ")
    ref = dense_tiles(golden, "integ")
    assert sorted(ps.list_tiles(0)) == sorted(ref.keys())
    for (cx, cy), a in ref.items():
        assert np.array_equal(tenths(ps.export_tile(0, cx, cy)), tenths(a))
    assert ps.stats()["cells_dropped"] == 0


def test_integration_clip_zero_range_and_degenerate_rays(PS, golden):
    ps = PS(2, 180, pool_subtiles=400)
    ps.poses = golden["clip_pose"]
    ps.set_scan(golden["clip_ranges"], golden["intel_angles"])
    ps.integrate()
    ref = dense_tiles(golden, "clip")
    assert sorted(ps.list_tiles(1)) == sorted(ref.keys())
    for (cx, cy), a in ref.items():
        assert np.array_equal(tenths(ps.export_tile(1, cx, cy)), tenths(a))
    # axis-aligned rays: pose on the lattice, beams along +-x / +-y (hybridmap.py:278-281)
    ang = np.array([0.0, np.pi / 2, np.pi, -np.pi / 2] * 45)
    rng = np.random.default_rng(3).uniform(0.5, 20.0, 180)
    ps2 = PS(1, 180, pool_subtiles=400)
    m = O.Map()
    for pose in ((0.0, 0.0, 0.0), (-1.0, 2.0, 0.0), (3.025, -4.025, np.pi / 2)):
        ps2.poses = pose
        ps2.set_scan(rng, ang)
        ps2.integrate()
        m.update(pose, O.Scan(rng, ang))
    assert_map_equal(ps2, 0, m, "axis-aligned")


def test_integration_per_particle_poses_and_saturation(PS, golden):
    """Every particle at its own pose; repeated integration drives cells into the
    +-3.0 clamps where the update order matters."""
    N = 24
    rng = np.random.default_rng(7)
    poses = np.column_stack([rng.uniform(-30, 30, N), rng.uniform(-30, 30, N), rng.uniform(-np.pi, np.pi, N)])
    poses[0] = (0, 0, 0)
    poses[1] = (19.99, 19.99, 0.3)          # next to a tile corner
    poses[2] = (-20.0, -20.0, 1.0)          # exactly on a tile border
    ps = PS(N, 180, pool_subtiles=2000)
    maps = [O.Map() for _ in range(N)]
    for rep in range(6):
        si = rep % 3
        ps.poses = poses
        ps.set_scan(*scan_of(golden, si))
        ps.integrate()
        for i in range(N):
            maps[i].update(poses[i], oscan(golden, si))
    ps.synchronize()
    for i in range(N):
        assert_map_equal(ps, i, maps[i], "particle %d" % i)


def test_integration_skipped_when_robot_in_no_tile(PS, golden):
    ps = PS(1, 180, pool_subtiles=64)
    ps.poses = (45.0, 3.0, 0.0)             # tile (40,0) does not exist yet -> hybridmap.py:98-100
    ps.set_scan(*scan_of(golden, 0))
    ps.integrate()
    assert ps.list_tiles(0) == [(0, 0)]
    assert np.count_nonzero(ps.export_tile(0, 0, 0)) == 0


# ----------------------------------------------------------------- weights --

def seeded(PS, G, N, pool=400):
    ps = PS(N, 180, pool_subtiles=pool)
    m = O.Map()
    for p, si in zip(G["upd_seed_poses"], G["upd_seed_scan_idx"]):
        ps.poses = p
        ps.set_scan(*scan_of(G, si))
        ps.integrate()
        m.update(p, oscan(G, si))
    return ps, m


def test_weight_stage_against_reference_map_update(PS, golden):
    """robot.py:73-115 with the matcher answer injected at the MATLAB seam; the proposal samples are the
    ones the unmodified reference drew with np.random.multivariate_normal (robot.py:81)."""
    N = 4
    ps, m = seeded(PS, golden, N)
    assert_map_equal(ps, 0, m, "seed")
    pose0 = np.array([0.6, 0.15, 0.1])
    ps.poses = pose0
    ps.covs = golden["upd_prior_cov"]
    ps.set_scan(*scan_of(golden, 4))
    ps.set_match(pose0 + golden["upd_match_corr"], golden["upd_match_cov"], 1)
    np.random.seed(int(golden["upd_seed"]))                  # the same NumPy call on the same stream reproduces them
    assert np.array_equal(np.random.multivariate_normal(pose0 + golden["upd_match_corr"], golden["upd_match_cov"], 30),
                          golden["upd_guesses"])
    ps.weight_guesses(np.tile(golden["upd_guesses"][None], (N, 1, 1)))
    ps.integrate(fallback_weights=True)
    pose, cov, w = ps.poses, ps.covs, ps.weights
    for i in range(N):
        assert np.allclose(pose[i], golden["upd_pose"], rtol=0, atol=POSE_ATOL)
        assert np.allclose(cov[i], golden["upd_cov"], rtol=REL_TOL, atol=1e-20)
        assert np.isclose(w[i], golden["upd_weight"], rtol=REL_TOL)
    ref = dense_tiles(golden, "upd")
    for (cx, cy), a in ref.items():
        assert np.array_equal(tenths(ps.export_tile(2, cx, cy)), tenths(a))


def test_nan_cov_fallback(PS, golden):
    """robot.py:73-78: failed match keeps the odometry pose, integrates, weight += 1 + sum L."""
    ps, m = seeded(PS, golden, 2)
    pose0 = np.array([0.6, 0.15, 0.1])
    ps.poses = pose0
    prior = golden["upd_prior_cov"]
    ps.covs = prior
    ps.set_scan(*scan_of(golden, 4))
    ps.set_match(pose0, np.full((3, 3), np.nan), 0)
    ps.weight(np.zeros((2, 30, 3)))
    ps.integrate(fallback_weights=True)
    assert np.array_equal(ps.poses[0], golden["bad_pose"])
    assert np.array_equal(ps.covs[1], prior)
    assert np.isclose(ps.weights[0], golden["bad_weight"], rtol=1e-12)
    ref = dense_tiles(golden, "bad")
    for (cx, cy), a in ref.items():
        assert np.array_equal(tenths(ps.export_tile(0, cx, cy)), tenths(a))


def test_weight_stage_random_matches_vs_oracle(PS, golden):
    N = 16
    ps, m = seeded(PS, golden, N)
    rng = np.random.default_rng(11)
    base = np.array([0.6, 0.15, 0.1])
    mposes = base + rng.normal(0, [0.05, 0.05, 0.02], (N, 3))
    A = rng.normal(0, 1, (N, 3, 3)) * np.array([0.03, 0.03, 0.004])[None, :, None]
    mcovs = A @ np.transpose(A, (0, 2, 1)) + np.diag([2e-4, 2e-4, 2e-6])
    mcovs[:, 0, 2] = mcovs[:, 2, 0] = mcovs[:, 1, 2] = mcovs[:, 2, 1] = 0.0      # matcher covariances are block diagonal
    valid = np.ones(N, dtype=np.int32)
    valid[[3, 9]] = 0
    z = rng.standard_normal((N, 30, 3))
    ps.poses = mposes - 0.01
    ps.set_scan(*scan_of(golden, 5))
    ps.set_match(mposes, mcovs, valid)
    w0 = ps.weights
    ps.weight(z)
    pose, cov, w = ps.poses, ps.covs, ps.weights
    s = oscan(golden, 5)
    for i in range(N):
        if not valid[i]:
            assert w[i] == w0[i]
            continue
        g, prs = O.propose(mposes[i], mcovs[i], z[i])
        ww = m.sample_weight(g, s, prs)
        op, oc, norm = O.moments(g, ww)
        assert np.array_equal(pose[i], op)                           # bit for bit: same exp / sin / cos, same order
        assert np.array_equal(cov[i], oc)
        assert w[i] == norm + w0[i]


# ----------------------------------------------------------------- matcher --

def check_match(ps, maps, poses, covs, s, idx):
    res = ps.match_result()
    for i in idx:
        rx, ry = O.pose_range(covs[i])
        o = maps[i].match(poses[i], s, rx, ry)
        b = res["best"][i]
        assert (int(b[0]), int(b[1]), int(b[2])) == o["best"], "particle %d: gpu %s oracle %s" % (i, b, o["best"])
        assert int(b[3]) == o["M"]
        assert bool(res["valid"][i]) == o["valid"]
        assert res["score"][i] == o["score"]
        assert np.array_equal(res["pose"][i], o["pose"])
        if o["valid"]:
            assert np.array_equal(res["cov"][i], o["cov"].reshape(3, 3))      # integer moments -> identical float64
        else:
            assert np.isnan(res["cov"][i]).all()


def test_match_against_oracle(PS, golden):
    """Same argmax (i, j, k), score, validity and covariance as the restated CPU
    matcher, on maps built from real Intel scans, for small and full windows."""
    N = 12
    ps, m = seeded(PS, golden, N)
    rng = np.random.default_rng(5)
    base = np.array([0.6, 0.15, 0.1])
    poses = base + rng.normal(0, [0.15, 0.15, 0.08], (N, 3))
    poses[0] = base
    covs = np.zeros((N, 3, 3))
    sig = rng.uniform(0.0, 0.01, (N, 2))            # window = clamp(120 sigma, 0.1, 0.7), robot.py:62-65
    covs[:, 0, 0] = sig[:, 0] ** 2
    covs[:, 1, 1] = sig[:, 1] ** 2
    covs[0] = 0.0                                   # smallest window (0.1 m)
    covs[1] = np.diag([1.0, 1.0, 1.0])              # full window (0.7 m)
    ps.poses = poses
    ps.covs = covs
    s = oscan(golden, 4)
    ps.set_scan(*scan_of(golden, 4))
    ps.scan_match()
    check_match(ps, [m] * N, poses, covs, s, range(N))
    # the translation score slice at the best rotation, cell by cell
    for i in (1, 5):
        rx, ry = O.pose_range(covs[i])
        o = m.match(poses[i], s, rx, ry)
        assert np.array_equal(ps.match_slice(i), o["slice"])


def _refine_case(PS, golden, N=16):
    ps, m = seeded(PS, golden, N)
    rng = np.random.default_rng(15)
    base = np.array([0.6, 0.15, 0.1])
    poses = base + rng.normal(0, [0.15, 0.15, 0.08], (N, 3))
    covs = np.zeros((N, 3, 3))
    covs[:, 0, 0] = covs[:, 1, 1] = rng.uniform(0.002, 0.01, N) ** 2
    covs[1] = np.diag([1.0, 1.0, 1.0])
    ps.poses = poses
    ps.covs = covs
    return ps, m, poses, covs


def test_match_ndt_refine_against_oracle(PS, golden):
    """NDT stage (matchScanCustom.m:32-50) after the grid search: same number of score
    evaluations, same accept decision, same refined pose and NDT score as the oracle;
    the covariance stays the grid stage's.  The arithmetic is shared polynomial
    exp/sin/cos in a fixed summation order, so the comparison is bit-exact."""
    N = 16
    ps, m, poses, covs = _refine_case(PS, golden, N)
    s = oscan(golden, 4)
    ps.set_scan(*scan_of(golden, 4))
    ps.set_refine(True)
    ps.scan_match()
    res = ps.match_result()
    old = O.set_refine(True)
    try:
        outs = [m.match(poses[i], s, *O.pose_range(covs[i])) for i in range(N)]
    finally:
        O.set_refine(old)
    grid = [m.match(poses[i], s, *O.pose_range(covs[i])) for i in range(N)]
    assert any(o["ndt_accepted"] for o in outs) and any(o["ndt_evals"] > 3 for o in outs)
    for i, (o, g) in enumerate(zip(outs, grid)):
        b = res["best"][i]
        assert (int(b[0]), int(b[1]), int(b[2])) == o["best"] and bool(res["valid"][i]) == o["valid"]
        assert int(res["ndt_evals"][i]) == o["ndt_evals"], "particle %d" % i
        assert bool(res["ndt_accepted"][i]) == o["ndt_accepted"]
        assert np.array_equal(res["pose"][i], o["pose"]), "particle %d: %r vs %r" % (i, res["pose"][i], o["pose"])
        assert res["score"][i] == o["score"]
        if o["valid"]:
            assert np.array_equal(res["cov"][i], g["cov"].reshape(3, 3))      # matchScanCustom.m:22 covariance of the grid stage
            if o["ndt_accepted"]:
                rx, ry = O.pose_range(covs[i])
                d = o["pose"] - poses[i]
                assert abs(d[0]) < rx and abs(d[1]) < ry and abs(d[2]) < np.pi / 6 and 2 * o["score"] > g["score"]
            else:
                assert np.array_equal(o["pose"], g["pose"]) and o["score"] == g["score"]
    # switching the stage off again restores the grid result
    ps.set_refine(False)
    ps.scan_match()
    check_match(ps, [m] * N, poses, covs, s, range(N))
    assert not ps.match_result()["ndt_evals"].any()


def test_match_adj_ndt_refine_against_oracle(PS, golden):
    """NDT stage on the scan-to-previous-scan variant: the curr points are raw endpoints
    (cos/sin of the guess heading come from two maths libraries), hence a tolerance."""
    N = 10
    rng = np.random.default_rng(8)
    ps = PS(N, 180, pool_subtiles=64, ndt_refine=True)
    prev_pose = np.array([0.5, 0.1, 0.2])
    gx, gy = O.transform(prev_pose, oscan(golden, 3))
    prev_xy = np.column_stack((gx, gy))
    base = np.array([0.55, 0.12, -0.25])
    poses = base + rng.normal(0, [0.1, 0.1, 0.05], (N, 3))
    covs = np.zeros((N, 3, 3))
    covs[:, 0, 0] = covs[:, 1, 1] = rng.uniform(0.002, 0.01, N) ** 2
    ps.poses = poses
    ps.covs = covs
    s = oscan(golden, 4)
    ps.set_scan(*scan_of(golden, 4))
    ps.scan_match(prev_xy)
    res = ps.match_result()
    old = O.set_refine(True)
    try:
        outs = [O.match_adj(poses[i], s, prev_xy, *O.pose_range(covs[i])) for i in range(N)]
    finally:
        O.set_refine(old)
    assert any(o["ndt_accepted"] for o in outs)
    for i, o in enumerate(outs):
        assert bool(res["valid"][i]) == o["valid"] and bool(res["ndt_accepted"][i]) == o["ndt_accepted"]
        assert np.allclose(res["pose"][i], o["pose"], rtol=0, atol=1e-7)
        assert np.isclose(res["score"][i], o["score"], rtol=1e-7)


def test_match_empty_map_is_invalid(PS, golden):
    """No occupied cell -> every score 0 -> best is the zero correction ->
    isValidPose rejects it (matchScanCustom.m:55) -> NaN covariance."""
    ps = PS(2, 180, pool_subtiles=64)
    ps.covs = np.diag([1.0, 1.0, 1.0])
    ps.set_scan(*scan_of(golden, 0))
    ps.scan_match()
    res = ps.match_result()
    assert not res["valid"].any()
    assert np.isnan(res["cov"]).all()
    assert (res["best"][:, :3] == 0).all()
    assert (res["score"] == 0).all()


# ---------------------------------------------------------------- resample --

def test_resample_golden_ancestors_bit_exact(PS, golden):
    for i in range(int(golden["rs_n"])):
        w = golden["rs%d_w" % i]
        ps = PS(len(w), 180, pool_subtiles=64)
        ps.weights = w
        did, anc = ps.resample(float(golden["rs%d_u" % i]))
        assert did == bool(golden["rs%d_did" % i])
        assert np.array_equal(anc, golden["rs%d_anc" % i])
        if did:
            assert (ps.weights == 1.0).all()                       # main.py:77-78
        else:
            assert np.array_equal(ps.weights, np.where(np.isinf(w), w, w))


@pytest.mark.parametrize("n", [1, 2, 31, 1000, 4096, 65536])
def test_resample_vs_oracle_sizes(PS, n):
    rng = np.random.default_rng(n)
    w = rng.normal(0, 1, n) * 1e6 - 3e5
    if n > 2:
        w[rng.integers(0, n, 3)] = -np.inf
    u = float(rng.random())
    ps = PS(n, 180, world_tiles=(1, 1), pool_subtiles=64)
    ps.weights = w
    poses = rng.normal(0, 1, (n, 3))
    ps.poses = poses
    did, anc = ps.resample(u)
    rc, oanc = O.resample(w, u)
    assert did == bool(rc == 1)
    assert np.array_equal(anc, oanc)
    assert np.all(np.diff(anc) >= 0) and len(anc) == n
    assert np.array_equal(ps.poses, poses[anc])                    # Robot.copy robot.py:141-149


def _cumsum_cases():
    rng = np.random.default_rng(77)
    n = 20000
    cases = {}
    cases["lognormal"] = np.exp(rng.normal(5.0, 3.0, n))
    cases["negative_shift"] = rng.normal(0, 1, n) * 1e6 - 3e5                 # main.py:54-55 path, some -inf
    cases["negative_shift"][rng.integers(0, n, 5)] = -np.inf
    w = np.ones(n)
    w[0] = 2.0 ** 53                                                         # every later add is a tie: 2^53 + 1 -> 2^53
    cases["all_ties"] = w
    w = rng.integers(1, 9, n).astype(np.float64) * 0.5                       # halves and integers: ties at every binade
    w[::97] = 1000.0
    cases["halves"] = w
    w = np.exp(rng.normal(0.0, 1.0, n))
    w[rng.integers(0, n, n // 3)] = 0.0                                      # zeros are skipped by the shift and add exactly
    w[7] = 500.0
    cases["zeros"] = w
    w = np.exp(rng.normal(0.0, 12.0, n))                                     # 30 binades of dynamic range
    cases["wide"] = w
    w = np.full(n, 5e-324)
    w[:3] = [1e-310, 300.0, 2.2250738585072014e-308]                         # denormals next to normal weights
    cases["denormals"] = w
    w = 2.0 ** rng.integers(-30, 11, n).astype(np.float64)                   # powers of two: exact halves of an ulp are common
    cases["powers_of_two"] = w
    return cases


@pytest.mark.parametrize("name", sorted(_cumsum_cases().keys()))
def test_resample_running_sum_bit_exact(PS, name):
    """The running sum of the weights is the one quantity of the path whose float64
    ROUNDING ORDER matters (main.py:57,62 adds left to right).  The kernel replaces the
    chain of N dependent adds by an exact integer prefix sum wherever the sum stays in one
    binade and no add is a tie; every c_i must equal the sequential sum bit for bit."""
    w = _cumsum_cases()[name]
    n = len(w)
    ps = PS(n, 180, world_tiles=(1, 1), pool_subtiles=64)
    ps.weights = w
    did, anc = ps.resample(0.618)
    assert did
    adj = np.where(np.isneginf(w), 0.0, w)
    mn = adj.min()
    if mn < 0:
        adj = np.where(adj != 0.0, adj + abs(mn), adj)                       # main.py:54-55
    ref = np.empty(n)
    c = 0.0
    for i, v in enumerate(adj.tolist()):                                     # plain Python floats: sequential float64
        c += v
        ref[i] = c
    got = ps.resample_cumsum()
    bad = np.flatnonzero(got.view(np.uint64) != ref.view(np.uint64))
    assert bad.size == 0, "first mismatch at %d: %r vs %r" % (bad[0], got[bad[0]], ref[bad[0]])
    rc, oanc = O.resample(w, 0.618)
    assert rc == 1 and np.array_equal(anc, oanc)


def test_resample_not_triggered_keeps_particles(PS):
    ps = PS(64, 180, world_tiles=(1, 1), pool_subtiles=64)
    w = np.linspace(1.0, 150.0, 64)                                # max - min <= 200, main.py:50
    ps.weights = w
    did, anc = ps.resample(0.3)
    assert not did and np.array_equal(anc, np.arange(64))
    assert np.array_equal(ps.weights, w)


def test_cow_duplicates_stay_independent(PS, golden):
    """After a resample duplicates share sub-tiles; integrating different poses
    must copy-on-write so that each particle matches its own deep-copied oracle map."""
    N = 8
    ps, m0 = seeded(PS, golden, N, pool=600)
    maps = [m0.copy() for _ in range(N)]
    w = np.array([1.0, 5000.0, 1.0, 1.0, 9000.0, 1.0, 1.0, 1.0])
    ps.weights = w
    did, anc = ps.resample(0.37)
    assert did
    maps = [maps[a].copy() for a in anc]
    st = ps.stats()
    assert st["shared_refs"] == st["total_refs"] > 0               # everything shared right after the copy
    used_before = st["pool_in_use"]
    rng = np.random.default_rng(2)
    poses = np.array([0.6, 0.15, 0.1]) + rng.normal(0, [0.3, 0.3, 0.2], (N, 3))
    for si in (5, 6):
        ps.poses = poses
        ps.set_scan(*scan_of(golden, si))
        ps.integrate()
        for i in range(N):
            maps[i].update(poses[i], oscan(golden, si))
    for i in range(N):
        assert_map_equal(ps, i, maps[i], "dup %d" % i)
    st = ps.stats()
    assert st["cow_copies"] > 0 and st["pool_in_use"] > used_before
    # dropping all but one lineage returns the other particles' private tiles to the pool
    ps.weights = np.array([9000.0] + [1.0] * (N - 1))
    did, anc = ps.resample(0.5)
    assert did and (anc == 0).all()
    assert ps.stats()["pool_in_use"] < st["pool_in_use"]
    assert_map_equal(ps, N - 1, maps[0], "after collapse")


# -------------------------------------------------------------- end to end --

def test_filter_steps_against_oracle(PS, golden):
    """main.py:152-160 for several scans: same normals and uniforms on both sides;
    ancestors bit-exact, poses/weights within tolerance, maps equal in tenths."""
    N, K, B = 6, 30, 180
    rng = np.random.default_rng(99)
    ps = PS(N, B, pool_subtiles=1500)
    f = O.Filter(N, B, K)
    par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
    r0, ang = scan_of(golden, 0)
    for _ in range(2):                                             # map seeding, main.py:89-90
        ps.set_scan(r0, ang); ps.integrate()
        f.set_scan(r0, ang); f.integrate()
    for step in range(1, 7):
        u = (rng.normal(0.05, 0.02), rng.normal(0.0, 0.02), rng.normal(-0.3, 0.05))
        ps.motion(1, u, 1.0, par)
        f.motion(1, u, 1.0, par)
        r, ang = scan_of(golden, step)
        z = rng.standard_normal((N, K, 3))
        u01 = float(rng.random())
        ps.set_scan(r, ang)
        ps.scan_match()
        ps.weight(z)
        ps.integrate(fallback_weights=True)
        f.set_scan(r, ang)
        f.map_update(z)
        res = ps.match_result()
        assert np.array_equal(res["valid"], f.valid.astype(bool)), "step %d" % step
        assert np.array_equal(ps.weights, f.weight), "step %d" % step
        did, anc = ps.resample(u01)
        odid, oanc = f.resample(u01)
        assert did == odid and np.array_equal(anc, oanc), "step %d" % step
        assert np.array_equal(ps.poses, f.pose), "step %d" % step
        assert np.array_equal(ps.covs, f.cov), "step %d" % step
    for i in range(N):
        assert_map_equal(ps, i, f.map(i), "final %d" % i)


def test_fused_step_runs_and_conserves_pool(PS, golden):
    """rbpf_step (device-side draws): identical particles stay identical only
    until sampling; pool accounting must stay consistent and no error flag set."""
    N = 512
    ps = PS(N, 180, pool_subtiles=40000, seed=1234)
    r0, ang = scan_of(golden, 0)
    for _ in range(2):
        ps.set_scan(r0, ang); ps.integrate()
    for step in range(1, 6):
        ps.motion(1, (0.05, 0.0, -0.3), 1.0, (0.002, 0.05, 0.01 * np.pi / 180, 0.05))
        ps.step(*scan_of(golden, step))
    ps.synchronize()
    st = ps.stats()
    assert st["pool_in_use"] <= st["pool_subtiles"] and st["cells_dropped"] == 0
    assert np.isfinite(ps.poses).all() and np.isfinite(ps.weights).all()
    assert st["total_refs"] >= st["pool_in_use"] > 0


def test_match_adj_against_oracle(PS, golden):
    """Scan-to-previous-scan variant (hybridmap.py:147-191): same argmax, score,
    validity, covariance as the oracle for every particle; unsnapped curr points,
    occupancy = the previous scan rasterised on the lattice."""
    N = 10
    rng = np.random.default_rng(8)
    ps = PS(N, 180, pool_subtiles=64)
    prev_pose = np.array([0.5, 0.1, 0.2])
    gx, gy = O.transform(prev_pose, oscan(golden, 3))
    prev_xy = np.column_stack((gx, gy))
    base = np.array([0.55, 0.12, -0.25])
    poses = base + rng.normal(0, [0.1, 0.1, 0.05], (N, 3))
    covs = np.zeros((N, 3, 3))
    covs[:, 0, 0] = covs[:, 1, 1] = rng.uniform(0, 0.01, N) ** 2
    covs[0] = np.diag([1.0, 1.0, 1.0])
    ps.poses = poses
    ps.covs = covs
    s = oscan(golden, 4)
    ps.set_scan(*scan_of(golden, 4))
    ps.scan_match(prev_xy)
    res = ps.match_result()
    for i in range(N):
        rx, ry = O.pose_range(covs[i])
        o = O.match_adj(poses[i], s, prev_xy, rx, ry)
        b = res["best"][i]
        assert (int(b[0]), int(b[1]), int(b[2])) == o["best"], "particle %d" % i
        assert int(b[3]) == o["M"] and bool(res["valid"][i]) == o["valid"] and res["score"][i] == o["score"]
        assert np.array_equal(res["pose"][i], o["pose"])
        if o["valid"]:
            assert np.array_equal(res["cov"][i], o["cov"])
    assert res["valid"].any()
