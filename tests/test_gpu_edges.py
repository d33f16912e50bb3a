"""GPU parity on the edge cases the domain offers: one particle, one beam, the
maximum beam count, odd beam counts (361 UNSW/Bele), all-zero and all-out-of-range
sweeps, other sample counts -- CUDA path vs oracle, same tolerances as the main
parity tests."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PS():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200.particles import ParticleSet

    return ParticleSet


def tenths(a):
    return np.rint(np.asarray(a) * 10.0).astype(np.int32)


def full_step_vs_oracle(PS, ranges_list, angles, N, K, steps_u, seed=0, pool=3000):
    """Seed with scan 0 twice, then one full update per remaining scan; compare everything."""
    B = len(angles)
    rng = np.random.default_rng(seed)
    ps = PS(N, B, n_samples=K, pool_subtiles=pool)
    f = O.Filter(N, B, K)
    for _ in range(2):
        ps.set_scan(ranges_list[0], angles); ps.integrate()
        f.set_scan(ranges_list[0], angles); f.integrate()
    par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
    for s in range(1, len(ranges_list)):
        ps.motion(1, steps_u, 1.0, par); f.motion(1, steps_u, 1.0, par)
        z = rng.standard_normal((N, K, 3))
        u01 = float(rng.random())
        ps.set_scan(ranges_list[s], angles); ps.scan_match(); ps.weight(z); ps.integrate(fallback_weights=True)
        f.set_scan(ranges_list[s], angles); f.map_update(z)
        assert np.array_equal(ps.match_result()["valid"], f.valid.astype(bool)), "scan %d" % s
        assert np.array_equal(ps.weights, f.weight), "scan %d" % s
        did, anc = ps.resample(u01)
        odid, oanc = f.resample(u01)
        assert did == odid and np.array_equal(anc, oanc), "scan %d" % s
        assert np.array_equal(ps.poses, f.pose), "scan %d" % s
    for i in range(N):
        ot = f.map(i).tiles()
        assert sorted(ps.list_tiles(i)) == sorted(ot.keys())
        for c, ref in ot.items():
            assert np.array_equal(tenths(ps.export_tile(i, *c)), tenths(ref)), "particle %d tile %s" % (i, c)
    return ps, f


def test_single_particle_single_gpu(PS, golden):
    r = [golden["intel_ranges"][i] for i in range(4)]
    full_step_vs_oracle(PS, r, golden["intel_angles"], N=1, K=30, steps_u=(0.05, 0.0, -0.45))


@pytest.mark.parametrize("K", [1, 8, 32])
def test_other_sample_counts(PS, golden, K):
    r = [golden["intel_ranges"][i] for i in range(3)]
    full_step_vs_oracle(PS, r, golden["intel_angles"], N=5, K=K, steps_u=(0.05, 0.0, -0.45), seed=K)


@pytest.mark.parametrize("B", [1, 2, 181, 361, 384])
def test_beam_counts(PS, B):
    """1 beam, 2 beams, 181 (the committed orebro.log), 361 (UNSW / Bele / CSAIL), 384 (the ABI maximum)."""
    from thesis_b200 import synth

    w = synth.Workload(4, n_beams=max(B, 2))
    ang = w.angles[:B]
    r = [w.ranges[i][:B] for i in range(4)]
    full_step_vs_oracle(PS, r, ang, N=3, K=30, steps_u=tuple(w.odom[0]), seed=B)


def test_degenerate_sweeps(PS, golden):
    """All-zero ranges (every beam marks the robot's own cell, no min-range gate in
    hybridmap.py:95-145), all ranges beyond every gate (no matcher points, no weight
    lookups, rays clipped at 15 m and left free), a sweep of NaN-free huge values."""
    ang = golden["intel_angles"]
    zero = np.zeros(180)
    far = np.full(180, 81.83)
    mixed = golden["intel_ranges"][1].copy()
    mixed[::3] = 0.0
    mixed[1::3] = 60.0
    for seq in ([golden["intel_ranges"][0], zero, far, golden["intel_ranges"][1]],
                [far, far, zero, zero],
                [golden["intel_ranges"][0], mixed, mixed]):
        full_step_vs_oracle(PS, seq, ang, N=3, K=30, steps_u=(0.02, 0.01, 0.1), seed=len(seq))


def test_rotated_lidar_mounts(PS, golden):
    """Beam fans that are not symmetric about the heading (full 360 degrees, reversed order)."""
    rng = np.random.default_rng(1)
    ang360 = np.linspace(-np.pi, np.pi, 360, endpoint=False)
    r = [np.clip(4.0 + rng.normal(0, 1.5, 360), 0.3, 20.0) for _ in range(3)]
    full_step_vs_oracle(PS, r, ang360, N=2, K=30, steps_u=(0.03, -0.02, 0.2))
    rev = golden["intel_angles"][::-1].copy()
    r2 = [golden["intel_ranges"][i][::-1].copy() for i in range(3)]
    full_step_vs_oracle(PS, r2, rev, N=2, K=30, steps_u=(0.05, 0.0, -0.45))
