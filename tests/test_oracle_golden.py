"""Pins oracle/rbpf_oracle.c against vectors produced by the Python reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle as O

DIM = 800


def dense_tiles(G, prefix):
    out = {}
    for n, (cx, cy) in enumerate(G[prefix + "_centres"]):
        a = np.zeros(DIM * DIM)
        a[G["%s_t%d_idx" % (prefix, n)]] = G["%s_t%d_val" % (prefix, n)]
        out[(int(cx), int(cy))] = a.reshape(DIM, DIM)
    return out


def assert_tiles_equal(omap, G, prefix):
    ref = dense_tiles(G, prefix)
    got = omap.tiles()
    # allocation order is part of the contract (hybridmap.py:131 appends)
    assert list(got.keys()) == [tuple(int(v) for v in c) for c in G[prefix + "_centres"]]
    for k in ref:
        assert np.array_equal(got[k], ref[k]), "tile %s differs" % (k,)


def scan(G, i):
    return O.Scan(G["intel_ranges"][int(i)], G["intel_angles"])


def test_beam_geometry_bit_exact(golden):
    s = scan(golden, 0)
    assert np.array_equal(np.stack([s.px, s.py]), golden["scan0_xy"])
    gx, gy = O.transform(golden["xform_pose"], scan(golden, 3))
    # bit-exact on the machine that produced the vectors (FMA order of np.matmul);
    # a different BLAS may differ in the last ulp
    assert np.allclose(np.stack([gx, gy]), golden["xform_xy"], rtol=0, atol=2e-14)


def test_bresenham_matches_reference(golden):
    rays, cells, offs = golden["bres_rays"], golden["bres_cells"], golden["bres_offs"]
    for n, (x0, y0, x1, y1) in enumerate(rays):
        got = O.bresenham(int(x0), int(y0), int(x1), int(y1))
        assert np.array_equal(got, cells[offs[n]:offs[n + 1]]), (x0, y0, x1, y1)
    # the documented quirk: negative axis-aligned rays are empty
    assert len(O.bresenham(5, 5, 5, 2)) == 0 and len(O.bresenham(5, 5, 2, 5)) == 0
    assert len(O.bresenham(3, 3, 3, 3)) == 1


def test_map_integration_bit_exact(golden):
    m = O.Map()
    for p, si in zip(golden["integ_poses"], golden["integ_scan_idx"]):
        m.update(p, scan(golden, si))
    assert_tiles_equal(m, golden, "integ")


def test_map_integration_clip_and_zero_range(golden):
    m = O.Map()
    m.update(golden["clip_pose"], O.Scan(golden["clip_ranges"], golden["intel_angles"]))
    assert_tiles_equal(m, golden, "clip")


def test_odds_lookup(golden):
    m = O.Map()
    for p, si in zip(golden["integ_poses"], golden["integ_scan_idx"]):
        m.update(p, scan(golden, si))
    for (x, y), v in zip(golden["odds_pts"], golden["odds_vals"]):
        got = m.odds_at(x, y)
        if np.isnan(v):
            assert got is None
        else:
            assert got == v


def seeded_map(G):
    m = O.Map()
    for p, si in zip(G["upd_seed_poses"], G["upd_seed_scan_idx"]):
        m.update(p, scan(G, si))
    return m


def test_sample_weight(golden):
    m = seeded_map(golden)
    w = m.sample_weight(golden["sw_guesses"], scan(golden, golden["sw_scan_idx"]), golden["sw_prs"])
    # reference accumulates in np.longdouble (robot.py:119,124); we pin float64
    assert np.allclose(w, golden["sw_w"], rtol=1e-13, atol=0)


def test_matcher_front_end_points(golden):
    """curr / ref point sets the reference hands to MATLAB (hybridmap.py:216-240)."""
    m = seeded_map(golden)
    guess = np.array([0.6, 0.15, 0.1])
    s = scan(golden, 4)
    curr = m.match_curr(guess, s)
    assert curr.shape == golden["upd_curr"].shape
    assert np.array_equal(curr, golden["upd_curr"])
    # ref = unique occupied cells within the 72x72 windows of the curr points, < 11.5 m: the set the restated
    # matcher scores on (orc_match masks the occupancy with it) is the reference's valid_ref_points, in np.unique order
    ref = m.match_ref(guess, s)
    assert ref.shape == golden["upd_ref"].shape
    assert np.array_equal(ref, golden["upd_ref"])
    pts = []
    for cx, cy in curr + guess[:2]:
        pts.append(m.nearby_occ(cx, cy))
    ref = np.unique(np.concatenate(pts), axis=0) - guess[:2]
    ref = ref[np.sqrt(ref[:, 0] ** 2 + ref[:, 1] ** 2) < 11.5]
    assert ref.shape == golden["upd_ref"].shape
    assert np.allclose(ref, golden["upd_ref"], rtol=0, atol=1e-12)


def test_full_map_update(golden):
    m = seeded_map(golden)
    s = scan(golden, 4)
    rx, ry = O.pose_range(golden["upd_prior_cov"])
    assert np.allclose([rx, ry], golden["upd_prange"][:2], rtol=0, atol=0)
    mean = np.array([0.6, 0.15, 0.1]) + golden["upd_match_corr"]
    np.random.seed(int(golden["upd_seed"]))                  # robot.py:81 with NumPy itself, same stream as the reference run
    g, prs = O.propose_numpy(mean, golden["upd_match_cov"], 30)
    assert np.array_equal(g, golden["upd_guesses"])
    w = m.sample_weight(g, s, prs)
    pose, cov, norm = O.moments(g, w)
    assert np.allclose(pose, golden["upd_pose"], rtol=0, atol=1e-13)
    assert np.allclose(cov, golden["upd_cov"], rtol=1e-12, atol=1e-20)
    assert np.isclose(norm + 1.0, golden["upd_weight"], rtol=1e-12)
    m.update(pose, s)
    assert_tiles_equal(m, golden, "upd")


def test_pdf_against_scipy(golden):
    import scipy.stats

    mean = np.array([0.62, 0.12, 0.11])
    z = np.random.default_rng(5).standard_normal((30, 3))
    g, prs = O.propose(mean, golden["upd_match_cov"], z)
    ref = scipy.stats.multivariate_normal.pdf(g, mean, golden["upd_match_cov"]) * 10   # robot.py:87
    assert np.allclose(prs, ref, rtol=1e-12)
    np.random.seed(3)
    g, prs = O.propose_numpy(mean, golden["upd_match_cov"], 30)
    ref = scipy.stats.multivariate_normal.pdf(g, mean, golden["upd_match_cov"]) * 10
    assert np.allclose(prs, ref, rtol=1e-12)


def test_nan_cov_fallback(golden):
    m = seeded_map(golden)
    s = scan(golden, 4)
    pose = np.array([0.6, 0.15, 0.1])
    m.update(pose, s)                                           # robot.py:75
    w = m.sample_weight(pose[None, :], s, np.array([1.0]))      # robot.py:76
    assert np.isclose(w[0] + 1.0, golden["bad_weight"], rtol=1e-13)
    assert np.array_equal(pose, golden["bad_pose"]) and int(golden["bad_nhist"]) == 2
    assert_tiles_equal(m, golden, "bad")


def test_pose_range(golden):
    for c, want in zip(golden["prange_covs"], golden["prange_out"]):
        assert np.array_equal(O.pose_range(c), want)


def test_resample_ancestors_bit_exact(golden):
    for i in range(int(golden["rs_n"])):
        rc, anc = O.resample(golden["rs%d_w" % i], float(golden["rs%d_u" % i]))
        assert bool(rc) == bool(golden["rs%d_did" % i])
        assert np.array_equal(anc, golden["rs%d_anc" % i])


@pytest.mark.parametrize("name,family,par", [
    ("abs", 0, (0, 0, 0, 0)),
    ("velraw", 1, (0.002, 0.05, 0.01 * np.pi / 180, 0.05)),
    ("velaces", 1, (0.02, 0.01, 0.2 * np.pi / 180, 0.02)),
    ("uni", 2, (0, 0, 0, 0)),
])
def test_motion_families(golden, name, family, par):
    dt = float(golden["mot_%s_dt_ticks" % name]) / 1e4
    pose, cov = O.motion(family, golden["mot_%s_u" % name], dt, par, golden["mot_pose0"], golden["mot_cov0"])
    assert np.allclose(pose, golden["mot_%s_pose" % name], rtol=0, atol=1e-14)
    assert np.allclose(cov, golden["mot_%s_cov" % name], rtol=1e-13, atol=1e-18)
