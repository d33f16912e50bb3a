"""GPU: the "next" rows of SURVEY 8f that ride on the hot path's data structures --
device-side occupied-point export (hybridmap.py:303-313) and checkpoint / resume of
the whole particle set (main.py:183-210 shelves particle 0 only)."""
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


def run_some(ps, G, steps, start=1):
    ang = G["intel_angles"]
    for s in range(start, start + steps):
        ps.motion(1, (0.05, 0.0, -0.4), 1.0, (0.002, 0.05, 0.01 * np.pi / 180, 0.05))
        ps.step(G["intel_ranges"][s], ang)


def test_occupied_points_equal_thresholded_tiles(P, golden):
    ps = P.ParticleSet(6, 180, pool_subtiles=1500, seed=2)
    ang = golden["intel_angles"]
    ps.set_scan(golden["intel_ranges"][0], ang); ps.integrate(); ps.integrate()
    run_some(ps, golden, 4)
    for p in (0, 5):
        got = ps.occupied_points(p)
        want = []
        for (cx, cy) in ps.list_tiles(p):
            t = ps.export_tile(p, cx, cy)
            ix, iy = np.nonzero(np.rint(t * 10.0) > 10)                    # OCCUPIED_POINT_THRESHOLD, hybridmap.py:310
            want.append(np.column_stack((((ix - 400) * 0.05 + cx) / 0.05, ((iy - 400) * 0.05 + cy) / 0.05)))
        want = np.concatenate(want)
        assert len(got) == len(want) > 100
        key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
        assert np.allclose(key(got), key(want), rtol=0, atol=1e-9)


def test_checkpoint_resume_continues_bit_identically(P, golden, tmp_path):
    """Save after 3 scans, keep going; a fresh handle that loads the checkpoint and
    replays the same scans with the same seed must end in the identical state."""
    mk = lambda: P.ParticleSet(48, 180, pool_subtiles=6000, seed=9)
    a = mk()
    ang = golden["intel_angles"]
    a.set_scan(golden["intel_ranges"][0], ang); a.integrate(); a.integrate()
    run_some(a, golden, 3)
    path = tmp_path / "rbpf.ckpt"
    a.save(path)
    st_a = a.stats()
    run_some(a, golden, 3, start=4)
    b = mk()
    b.load(path)
    st_b = b.stats()
    for k in ("pool_in_use", "total_refs", "shared_refs", "refcount_sum"):
        assert st_a[k] == st_b[k], k
    run_some(b, golden, 3, start=4)
    assert np.array_equal(a.poses, b.poses) and np.array_equal(a.weights, b.weights) and np.array_equal(a.covs, b.covs)
    for p in (0, 17, 47):
        assert a.list_tiles(p) == b.list_tiles(p)
        for c in a.list_tiles(p):
            assert np.array_equal(a.export_tile(p, *c), b.export_tile(p, *c))
    # a handle with another configuration refuses the file
    c = P.ParticleSet(47, 180, pool_subtiles=6000)
    with pytest.raises(P.RbpfError):
        c.load(path)
