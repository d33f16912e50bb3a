"""GPU, slow: parity against the oracle at configuration size and over a whole log.

  * configs[2] (ACES-shaped stand-in, 180 beams): 8,192 particles x 3 scans against the oracle
    outright -- weights, poses and covariances bit for bit, ancestors, maps of sampled particles.
  * the complete Intel Research Lab log (tests/golden/intel.txt.gz, the reference's data/intel.txt,
    903 sweeps after timestamp de-duplication) through the headless main.py loop with main.py's
    map / adj alternation: the CUDA path and the oracle in lockstep on the same draws; the
    weights must be equal bit for bit before EVERY resample and the ancestors equal after it.
    RBPF_FULL_LOG_PARTICLES sets the particle count (default 256: about 2.5 minutes of oracle time
    on 16 host cores; profiles/ holds the log of a 1,024-particle run).
"""
import gzip
import os
import shutil
import sys
import time

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_ROOT, os.path.join(_ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def P():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200 import particles

    return particles


def tenths(a):
    return np.rint(np.asarray(a) * 10).astype(np.int64)


def assert_maps_equal(ps, f, particles):
    for i in particles:
        ot = f.map(int(i)).tiles()
        assert sorted(ps.list_tiles(int(i))) == sorted(ot.keys()), "tile set of particle %d" % i
        for c, ref in ot.items():
            assert np.array_equal(tenths(ps.export_tile(int(i), *c)), tenths(ref)), "particle %d tile %s" % (i, c)


def test_config3_aces_shape_8192_particles_vs_oracle(P):
    """configs[2]: 8,192 particles, 180-beam sweeps, 5 cm cells (synthetic stand-in for the missing
    aces.txt) -- three scans with resampling against the oracle outright."""
    from thesis_b200 import synth

    N, K, B = 8192, 30, 180
    w = synth.Workload(5, B)
    rng = np.random.default_rng(21)
    ps = P.ParticleSet(N, B, world_tiles=(5, 5), pool_subtiles=30 * N)
    f = O.Filter(N, B, K)
    for _ in range(2):
        ps.set_scan(w.ranges[0], w.angles); ps.integrate()
        f.set_scan(w.ranges[0], w.angles); f.integrate()
    n_resampled = 0
    for s in range(1, 4):
        ps.motion(1, w.odom[s - 1], w.dt, w.par); f.motion(1, w.odom[s - 1], w.dt, w.par)
        z = rng.standard_normal((N, K, 3))
        u01 = float(rng.random())
        ps.set_scan(w.ranges[s], w.angles); ps.scan_match(); ps.weight(z); ps.integrate(fallback_weights=True)
        f.set_scan(w.ranges[s], w.angles); f.map_update(z)
        assert np.array_equal(ps.match_result()["valid"], f.valid.astype(bool)), "scan %d" % s
        assert np.array_equal(ps.weights, f.weight), "scan %d: weights must be bit-identical" % s
        did, anc = ps.resample(u01)
        odid, oanc = f.resample(u01)
        assert did == odid and np.array_equal(anc, oanc), "scan %d" % s
        n_resampled += int(did)
        assert np.array_equal(ps.poses, f.pose) and np.array_equal(ps.covs, f.cov), "scan %d" % s
    assert n_resampled >= 1
    assert_maps_equal(ps, f, (0, 1, 4095, 8191) + tuple(rng.integers(0, N, 4)))
    st = ps.stats()
    assert st["refcount_sum"] == st["total_refs"]


class _Lockstep:
    """One Robot-like view that drives the CUDA set and the oracle filter together (the harness
    calls only the first view of a device-backed list, thesis_b200/harness.py _fan)."""

    def __init__(self, ps, f, rng):
        self.ps, self.f, self.rng = ps, f, rng
        self._shared = self                       # harness: "device-backed", one call per round
        self.resamples = self.triggered = self.updates = 0
        self.bad_matches = 0

    # -- particle API used by harness.run_log
    def imu_update(self, reading):
        fam, par = reading.motion
        u = np.asarray(reading.get_data(), dtype=np.float64)
        self.ps.motion(fam, u, reading.dt() / 1e4, par)
        self.f.motion(fam, u, reading.dt() / 1e4, par)

    def map_update(self, scan, last_scan, adj):
        ps, f = self.ps, self.f
        z = self.rng.standard_normal((ps.N, ps.K, 3))
        prev = np.column_stack((last_scan.x(), last_scan.y())) if adj else None
        ps.set_scan(scan.ranges(), scan.angles())
        ps.scan_match(prev)
        ps.weight(z)
        ps.integrate(fallback_weights=True)
        f.set_scan(scan.ranges(), scan.angles())
        f.map_update(z, prev)
        self.updates += 1
        self.bad_matches += int(ps.N - ps.match_result()["valid"].sum())

    def get_latest_pose(self):
        from thesis_b200.models import Pose

        p = self.ps.poses[0]
        return Pose(float(p[0]), float(p[1]), float(p[2]))

    def resample(self, particles):
        ps, f = self.ps, self.f
        k = self.resamples
        assert np.array_equal(ps.weights, f.weight), "weights differ before resample %d" % k
        u01 = float(self.rng.random())
        did, anc = ps.resample(u01)
        odid, oanc = f.resample(u01)
        assert did == odid, "trigger differs at resample %d" % k
        assert np.array_equal(anc, oanc), "ancestors differ at resample %d" % k
        assert np.array_equal(ps.poses, f.pose), "poses differ after resample %d" % k
        self.resamples += 1
        self.triggered += int(did)
        return particles


def run_full_intel_log(P, n, tmp_dir, max_frames=None):
    from thesis_b200 import harness, loaders, sensors

    with gzip.open(os.path.join(GOLDEN, "intel.txt.gz"), "rb") as src, open(os.path.join(tmp_dir, "intel.txt"), "wb") as dst:
        shutil.copyfileobj(src, dst)

    class Lidar(loaders.IntelLidarData):
        FILE = "intel.txt"

        def __init__(self):
            super().__init__(tmp_dir)

    class Imu(loaders.IntelIMUData):
        FILE = "intel.txt"

        def __init__(self):
            super().__init__(tmp_dir)

    ld, im = sensors.Lidar(Lidar()), sensors.IMU(Imu())
    ps = P.ParticleSet(n, 180, world_tiles=(5, 5), pool_subtiles=max(4096, 130 * n))
    f = O.Filter(n, 180, 30)
    view = _Lockstep(ps, f, np.random.default_rng(2024))

    def seed(particles, scan):
        for _ in range(2):
            ps.set_scan(scan.ranges(), scan.angles()); ps.integrate()
            f.set_scan(scan.ranges(), scan.angles()); f.integrate()

    t0 = time.time()
    _, log = harness.run_log([view], ld, im, view.resample, seed_fn=seed, max_frames=max_frames)
    ps.synchronize()
    assert np.array_equal(ps.weights, f.weight) and np.array_equal(ps.covs, f.cov)
    assert_maps_equal(ps, f, sorted({0, n // 2, n - 1}))
    st = ps.stats()
    assert st["refcount_sum"] == st["total_refs"]
    return dict(particles=n, frames=len(log), updates=view.updates, resamples=view.resamples, triggered=view.triggered,
                failed_matches=view.bad_matches, seconds=time.time() - t0, pool_in_use=st["pool_in_use"],
                cells_dropped=st["cells_dropped"])


def test_full_intel_log_every_resample_vs_oracle(P, tmp_path):
    n = int(os.environ.get("RBPF_FULL_LOG_PARTICLES", "256"))
    r = run_full_intel_log(P, n, str(tmp_path))
    print("full Intel log:", r)
    assert r["frames"] >= 900 and r["updates"] >= 900 and r["resamples"] == r["updates"]
    assert r["triggered"] >= 100                   # the comparison is not vacuous


if __name__ == "__main__":                         # python tests/test_gpu_long.py 1024 -> one JSON line (profiles/)
    import json
    import tempfile

    from thesis_b200 import particles as _P

    with tempfile.TemporaryDirectory() as d:
        print(json.dumps(run_full_intel_log(_P, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, d)))
