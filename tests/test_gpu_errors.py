"""GPU: error conventions of the C ABI (include/rbpf_b200.h) -- status codes, not
crashes; a failed match is a flag, not an error (matchScanCustom.m:26-28)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200 import particles

    return particles


def test_pool_exhaustion_is_reported_not_fatal(P, golden):
    ps = P.ParticleSet(64, 180, pool_subtiles=40)            # far too small for 64 diverging particles
    r, a = golden["intel_ranges"][0], golden["intel_angles"]
    ps.poses = np.random.default_rng(0).uniform(-30, 30, (64, 3))
    ps.set_scan(r, a)
    ps.integrate()
    with pytest.raises(P.RbpfError, match="pool"):
        ps.synchronize()
    st = ps.stats()
    assert st["pool_in_use"] <= st["pool_subtiles"]


def test_rays_leaving_the_world_are_dropped_and_counted(P, golden):
    ps = P.ParticleSet(2, 180, world_tiles=(1, 1), pool_subtiles=200)     # 40 m x 40 m world
    ps.poses = (18.0, 0.0, 0.0)                                          # 2 m from the border, looking out
    ps.set_scan(np.full(180, 9.0), golden["intel_angles"])
    ps.integrate()
    ps.synchronize()
    st = ps.stats()
    assert st["cells_dropped"] > 0
    assert ps.list_tiles(0) == [(0, 0)]


def test_bad_arguments(P, golden):
    with pytest.raises(P.RbpfError):
        P.ParticleSet(4, 180, world_tiles=(4, 4))                         # even tile counts are rejected
    with pytest.raises(P.RbpfError):
        P.ParticleSet(4, 1000)                                            # more beams than RB_MAXB
    ps = P.ParticleSet(4, 180, pool_subtiles=64)
    with pytest.raises(P.RbpfError):
        ps.scan_match()                                                   # no scan set
    with pytest.raises(ValueError):
        ps.set_scan(golden["intel_ranges"][0], golden["intel_angles"])
        ps.weight(np.zeros(7))                                            # wrong number of normals
    assert ps.export_tile(0, 40, 0) is None                               # tile does not exist: None like HybridMap
    with pytest.raises(P.RbpfError):
        ps.export_tile(0, 41, 0)                                          # not a tile centre


def test_resample_assertion_maps_to_python_assertion(P):
    """main.py:66-67: AssertionError("Incorrect number of resampled weights.") -- reachable with NaN weights."""
    ps = P.ParticleSet(16, 180, world_tiles=(1, 1), pool_subtiles=64)
    w = np.linspace(0.0, 1000.0, 16)
    w[3] = np.nan
    ps.weights = w
    poses = np.arange(48, dtype=float).reshape(16, 3)
    ps.poses = poses
    try:
        did, anc = ps.resample(0.5)
    except AssertionError as e:
        assert "Incorrect number of resampled weights" in str(e)
        assert np.array_equal(ps.poses, poses)                            # particles unchanged
    else:
        # NaN compares false everywhere: max - min > 200 may not fire; then nothing must have changed
        assert np.array_equal(ps.poses, poses if not did else poses[anc])


def test_peer_mapping_errors(P):
    """Pull migration (include/rbpf_b200.h rbpf_peer_*): a peer whose layout differs, a
    second attach, an unattached peer and the handle's own rank are refused with a status."""
    from thesis_b200 import dist as D

    a = D.MigratingSet(8, 180, 0, 2, pool_subtiles=300)
    b = D.MigratingSet(8, 180, 1, 2, pool_subtiles=300)
    other_pool = D.MigratingSet(8, 180, 1, 2, pool_subtiles=200)
    other_n = D.MigratingSet(16, 180, 1, 2, pool_subtiles=300)
    with pytest.raises(P.RbpfError):
        a._ck(a._lib.rbpf_migrate_pull(a._h))                 # nobody attached yet
    with pytest.raises(P.RbpfError, match="differs"):
        a.attach_peer(1, other_pool.peer_view())
    with pytest.raises(P.RbpfError, match="differs"):
        a.attach_peer(1, other_n.peer_view())
    with pytest.raises(P.RbpfError):
        a.attach_peer(0, b.peer_view())                       # own rank
    with pytest.raises(P.RbpfError):
        a.attach_peer(2, b.peer_view())                       # outside the world
    a.attach_peer(1, b.peer_view())
    with pytest.raises(P.RbpfError, match="already"):
        a.attach_peer(1, b.peer_view())
    # buffers flipped on one side only: the parities disagree and the attach is refused
    c = D.MigratingSet(8, 180, 0, 2, pool_subtiles=300)
    b._ck(b._lib.rbpf_resample_commit(b._h))
    with pytest.raises(P.RbpfError, match="differs"):
        c.attach_peer(1, b.peer_view())


def test_pool_exhaustion_inside_step_keeps_the_pool_consistent(P, golden):
    """rbpf_step never synchronises: when the pool runs out inside a step the ray-cast is skipped,
    resampling is frozen, the free-list counter stays inside [0, pool] (resample_refs pushes with it)
    and the status surfaces from rbpf_step two steps later, sticky until rbpf_clear_errors."""
    N = 64
    ps = P.ParticleSet(N, 180, pool_subtiles=1200, seed=3)
    r, a = golden["intel_ranges"], golden["intel_angles"]
    ps.set_scan(r[0], a)
    ps.integrate()
    ps.integrate()
    ps.synchronize()
    rng = np.random.default_rng(1)
    seen = None
    for s in range(1, 16):
        # teleport inside the existing tile: every scan needs fresh private sub-tiles, and weights that resample
        ps.poses = np.column_stack([rng.uniform(-18, 18, N), rng.uniform(-18, 18, N), rng.uniform(-3, 3, N)])   # inside tile (0, 0)
        ps.weights = np.linspace(0.0, 1000.0, N)
        try:
            ps.step(r[s], a)
        except P.RbpfError as e:
            seen = (s, str(e))
            break
    assert seen is not None and "pool" in seen[1], "pool exhaustion must surface from rbpf_step itself"
    with pytest.raises(P.RbpfError, match="pool"):
        ps.synchronize()                                                  # sticky
    st = ps.stats()
    assert st["pool_in_use"] <= st["pool_subtiles"]
    assert st["refcount_sum"] == st["total_refs"]
    ps.clear_errors()
    ps.synchronize()                                                      # clean again
    did, anc = ps.resample(0.25)                                          # pushes freed sub-tiles through the counter
    st = ps.stats()
    assert st["refcount_sum"] == st["total_refs"]
    assert st["pool_in_use"] <= st["pool_subtiles"]


def test_resample_assertion_is_sticky_without_outputs(P):
    """rbpf_resample(h, u, NULL, NULL) cannot report: the assertion (main.py:66-67) is kept and
    returned by the next synchronising call."""
    import ctypes as C

    ps = P.ParticleSet(16, 180, world_tiles=(1, 1), pool_subtiles=64)
    w = np.linspace(0.0, 1000.0, 16)
    w[5] = np.inf                                                         # sum = inf: floor((c - u)/slice) is NaN
    ps.weights = w
    u = C.c_double(0.5)
    rc = ps._lib.rbpf_resample(ps._h, C.byref(u), None, None)
    assert rc == 0
    with pytest.raises(AssertionError, match="Incorrect number of resampled weights"):
        ps.synchronize()
    ps.clear_errors()
    ps.synchronize()


def test_checkpoint_read_validates(P, golden, tmp_path):
    ps = P.ParticleSet(8, 180, pool_subtiles=400)
    ps.set_scan(golden["intel_ranges"][0], golden["intel_angles"])
    ps.integrate()
    path = str(tmp_path / "ck.bin")
    ps.save(path)
    poses = ps.poses.copy()
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])                        # truncated
    ps.poses = poses + 1.0
    with pytest.raises(P.RbpfError, match="size"):
        ps.load(path)
    assert np.array_equal(ps.poses, poses + 1.0)                          # nothing was touched
    other = P.ParticleSet(8, 180, pool_subtiles=400, rank=1, world=2)
    open(path, "wb").write(data)
    with pytest.raises(P.RbpfError, match="rank"):
        other.load(path)
    ps.load(path)
    assert np.array_equal(ps.poses, poses)
