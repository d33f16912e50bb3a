"""GPU: the reference's particle API (Robot views, resample) over the CUDA path,
driven by the headless main.py loop on the Intel excerpt, against the run of the
reference's own Python classes (golden e2e_*) and against the oracle."""
import numpy as np
import pytest

from excerpt import ExcerptIMU, ExcerptLidar

pytestmark = pytest.mark.gpu
DIM = 800


@pytest.fixture(scope="module")
def gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200 import particles

    return particles


def test_drop_in_loop_matches_reference_run(gpu, golden):
    from thesis_b200 import harness, sensors

    n, frames = int(golden["e2e_n"]), int(golden["e2e_frames"])
    np.random.seed(0)
    ld, im = sensors.Lidar(ExcerptLidar()), sensors.IMU(ExcerptIMU())
    gpu.new_filter(rng="numpy", pool_subtiles=2000)
    parts = [gpu.Robot("eng") for _ in range(n)]               # main.py:87 verbatim
    anc, poses = [], []

    def resample(ps):
        out = gpu.resample(ps)
        anc.append(ps[0]._shared.last_ancestors.copy())
        return out

    parts, log = harness.run_log(parts, ld, im, resample, seed_fn=gpu.seed_map, max_frames=frames,
                                 on_frame=lambda f, ps: poses.append([list(p.get_latest_pose().as_tuple()) for p in ps]))
    assert np.array_equal([l["updated"] for l in log], golden["e2e_updated"])
    assert np.array_equal(np.array(anc), golden["e2e_ancestors"])                 # ancestor indices bit-exact
    assert np.allclose(np.array(poses), golden["e2e_poses"], rtol=0, atol=1e-9)   # stated pose tolerance
    w = np.array([p.weight()[-1] for p in parts])
    assert np.allclose(w, golden["e2e_weights"], rtol=1e-9)                       # stated weight tolerance
    ps = parts[0]._shared.ps
    for i in range(2):
        cen = [tuple(int(v) for v in c) for c in golden["e2e_p%d_centres" % i]]
        assert sorted(ps.list_tiles(i)) == sorted(cen)
        for k, c in enumerate(cen):
            a = np.zeros(DIM * DIM)
            a[golden["e2e_p%d_t%d_idx" % (i, k)]] = golden["e2e_p%d_t%d_val" % (i, k)]
            got = ps.export_tile(i, c[0], c[1])
            assert np.array_equal(np.rint(got * 10), np.rint(a.reshape(DIM, DIM) * 10))   # per-cell log-odds, exact in tenths
            assert np.max(np.abs(got - a.reshape(DIM, DIM))) < 1.2e-14
    # histories the reference plots (main.py:173)
    assert len(parts[0].x()) == len(parts[0].y()) == len(parts[0].theta()) > frames
    assert parts[0].get_latest_pose().x() == parts[0].x()[-1]


def test_map_view_occupied_points(gpu, golden):
    from thesis_b200 import sensors

    parts = gpu.make_particles(2, pool_subtiles=400)
    s = sensors.Scan(golden["intel_ranges"][0], golden["intel_angles"])
    gpu.seed_map(parts, s)
    xs, ys = parts[1]._map.get_occupied_points()
    t = parts[1]._shared.ps.export_tile(1, 0, 0)
    assert len(xs) == int(np.count_nonzero(np.rint(t * 10) > 10)) > 50
    assert parts[0]._map._cell_size == 0.05


def test_device_rng_mode_runs(gpu, golden):
    from thesis_b200 import harness, sensors

    ld, im = sensors.Lidar(ExcerptLidar()), sensors.IMU(ExcerptIMU())
    parts = gpu.make_particles(64, rng="device", keep_history=False, pool_subtiles=8000, seed=3)
    parts, log = harness.run_log(parts, ld, im, gpu.resample, seed_fn=gpu.seed_map, max_frames=8)
    ps = parts[0]._shared.ps
    ps.synchronize()
    assert np.isfinite(ps.poses).all() and ps.stats()["cells_dropped"] == 0


@pytest.mark.parametrize("refine", [False, True])
def test_long_run_gpu_vs_oracle_all_excerpt_frames(gpu, refine):
    """All 45 sweeps of the Intel excerpt, 48 particles, main.py's map/adj alternation:
    the GPU views and the oracle, driven by the same loop and the same NumPy draws,
    must pick the same ancestors at every resample and stay within tolerance.
    refine: with the NDT stage of the matcher (matchScanCustom.m:32-50) on in both."""
    import oracle as O
    import ref_adapter as RA
    from thesis_b200 import harness, sensors

    n = 48
    np.random.seed(7)
    ld, im = sensors.Lidar(ExcerptLidar()), sensors.IMU(ExcerptIMU())
    op = RA.OracleParticles(n, 180)
    oposes = []
    old = O.set_refine(refine)
    try:
        harness.run_log(op.views, ld, im, op.resample, seed_fn=op.seed,
                        on_frame=lambda f, ps: oposes.append([list(p.get_latest_pose().as_tuple()) for p in ps]))
    finally:
        O.set_refine(old)
    np.random.seed(7)
    ld, im = sensors.Lidar(ExcerptLidar()), sensors.IMU(ExcerptIMU())
    gpu.new_filter(rng="numpy", keep_history=False, pool_subtiles=20000, ndt_refine=refine)
    parts = [gpu.Robot("eng") for _ in range(n)]
    anc, gposes = [], []

    def resample(ps):
        out = gpu.resample(ps)
        anc.append(ps[0]._shared.last_ancestors.copy())
        return out

    parts, log = harness.run_log(parts, ld, im, resample, seed_fn=gpu.seed_map,
                                 on_frame=lambda f, ps: gposes.append(ps[0]._shared.poses().copy()))
    assert len(log) == 45 and len(anc) == len(op.ancestors) >= 40
    for k, (a, b) in enumerate(zip(anc, op.ancestors)):
        assert np.array_equal(a, b), "resample %d" % k
    assert np.allclose(np.array(gposes), np.array(oposes), rtol=0, atol=1e-8)
    ps = parts[0]._shared.ps
    assert np.allclose(ps.weights, op.f.weight, rtol=1e-8)
    for i in (0, 23, 47):
        ot = op.f.map(i).tiles()
        assert sorted(ps.list_tiles(i)) == sorted(ot.keys())
        for c, ref in ot.items():
            assert np.array_equal(np.rint(ps.export_tile(i, *c) * 10), np.rint(ref * 10))
