"""Test-only adapters that let the reference's own Robot class and the CPU
oracle run under thesis_b200.harness.run_log with the reference particle API.

RefParticles   -- the reference's robot.Robot (imported through oracle/ref_shim)
                  with the MATLAB engine replaced by the oracle's restated matcher
                  (run on a shadow oracle map kept in lock-step with the Python
                  HybridMap) and np.random.multivariate_normal replaced by our
                  declared sampling transform on np.random.standard_normal draws.
OracleParticles -- the oracle Filter behind Robot-like views.
"""
import contextlib
import io

import numpy as np

import oracle as O


class _Engine:
    """Stands in for the MATLAB engine of ONE robot (hybridmap.py:244-251)."""

    def __init__(self):
        self.shadow = O.Map()
        self.guess = None
        self.scan = None
        self.prange = None
        self.prev_xy = None            # set on adj frames (hybridmap.py:147-191)

    def matchScanCustom(self, curr, ref, guess0, res, prange, nargout=3):
        if self.prev_xy is not None:
            r = O.match_adj(self.guess, self.scan, self.prev_xy, float(prange[0]), float(prange[1]))
            assert r["M"] == len(curr), "oracle and reference disagree on valid_curr_points"
        else:
            r = self.shadow.match(self.guess, self.scan, float(prange[0]), float(prange[1]))
        self.last = r
        corr = r["pose"] - self.guess
        return [list(corr)], r["cov"].tolist(), r["score"]


def make_ref_particles(n):
    import ref_shim

    ref_shim.install()
    out = []
    for _ in range(n):
        r = ref_shim.fresh_robot()
        eng = _Engine()
        r._map._matlab = eng
        r._eng = eng
        out.append(r)
    return out


def ref_map_update(robot, scan, last_scan, adj):
    """robot.map_update with the MATLAB seam answered by the restated matcher; everything else,
    the sampling with np.random.multivariate_normal included, is the reference's own code."""
    import models

    eng = robot._eng
    eng.guess = np.array([robot._x[-1], robot._y[-1], robot._theta[-1]], dtype=np.float64)
    eng.scan = O.Scan(scan.ranges(), scan.angles())
    eng.prev_xy = np.column_stack((last_scan.x(), last_scan.y())) if adj else None
    # robot.py:59-115 unmodified: the real np.random.multivariate_normal draws the samples (robot.py:81)
    with contextlib.redirect_stdout(io.StringIO()):
        robot.map_update(scan, last_scan, bool(adj))
    robot._cov = np.array(robot._cov, dtype=np.float64)
    pose = np.array([robot._x[-1], robot._y[-1], robot._theta[-1]], dtype=np.float64)
    eng.shadow.update(pose, eng.scan)            # keep the shadow map in lock-step (hybridmap.py:95-145)


class RefParticle:
    """Wraps a reference Robot so that harness.run_log can call it."""

    def __init__(self, robot):
        self.r = robot

    def imu_update(self, reading):
        with contextlib.redirect_stdout(io.StringIO()):
            return self.r.imu_update(reading)

    def map_update(self, scan, last_scan, adj):
        ref_map_update(self.r, scan, last_scan, adj)

    def get_latest_pose(self):
        return self.r.get_latest_pose()

    def weight(self):
        return [float(w) for w in self.r._weight]


def ref_seed(particles, scan):
    import models

    for p in particles:
        s = O.Scan(scan.ranges(), scan.angles())
        for _ in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                p.r._map.update(p.r.get_latest_pose(), scan)
            p.r._eng.shadow.update(np.array([p.r._x[-1], p.r._y[-1], p.r._theta[-1]], dtype=np.float64), s)


def ref_resample(particles, ancestors_log=None):
    """main.resample on float64 weights; reference copies get a copied shadow map."""
    import main as refmain
    import robot as refrobot

    robots = [p.r for p in particles]
    for i, r in enumerate(robots):
        r._tag = i
        r._weight = [np.float64(w) for w in r._weight]
    orig_copy = refrobot.Robot.copy

    def copy(self):
        c = orig_copy(self)
        c._map._maps = list(c._map._maps)
        eng = _Engine()
        eng.shadow = self._eng.shadow.copy()
        c._map._matlab = eng
        c._eng = eng
        c._tag = self._tag
        return c

    refrobot.Robot.copy = copy
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            new = refmain.resample(robots)
    finally:
        refrobot.Robot.copy = orig_copy
    if ancestors_log is not None:
        ancestors_log.append(np.array([r._tag for r in new], dtype=np.int32))
    return [RefParticle(r) for r in new]


class OracleParticles:
    """The oracle Filter behind the reference particle API (one shared filter)."""

    class View:
        def __init__(self, owner, slot):
            self.o, self.slot = owner, slot
            self.mseen = self.useen = 0

        def imu_update(self, reading):
            o = self.o
            if self.mseen == o.mrounds:
                fam, par = reading.motion
                o.f.motion(fam, np.asarray(reading.get_data(), dtype=np.float64), reading.dt() / 1e4, par)
                o.mrounds += 1
            self.mseen += 1
            return self.get_latest_pose()

        def map_update(self, scan, last_scan, adj):
            o = self.o
            if self.useen == o.urounds:
                f = o.f
                f.set_scan(scan.ranges(), scan.angles())
                # the reference draws normals only for particles whose match is valid, in particle
                # order; the oracle filter decides validity inside map_update, so pre-compute it
                s = O.Scan(scan.ranges(), scan.angles())
                prev = np.column_stack((last_scan.x(), last_scan.y())) if adj else None
                g = np.zeros((f.N, f.K, 3))
                for i in range(f.N):
                    rx, ry = O.pose_range(f.cov[i])
                    r = O.match_adj(f.pose[i], s, prev, rx, ry) if adj else f.map(i).match(f.pose[i], s, rx, ry)
                    if r["valid"]:
                        g[i] = np.random.multivariate_normal(r["pose"], r["cov"], f.K)      # robot.py:81
                f.map_update(g, prev, guesses=True)
                o.urounds += 1
            self.useen += 1

        def get_latest_pose(self):
            from thesis_b200.models import Pose

            p = self.o.f.pose[self.slot]
            return Pose(float(p[0]), float(p[1]), float(p[2]))

        def weight(self):
            return [float(self.o.f.weight[self.slot])]

    def __init__(self, n, n_beams, K=30):
        self.f = O.Filter(n, n_beams, K)
        self.mrounds = self.urounds = 0
        self.views = [OracleParticles.View(self, i) for i in range(n)]
        self.ancestors = []

    def seed(self, particles, scan):
        self.f.set_scan(scan.ranges(), scan.angles())
        self.f.integrate()
        self.f.integrate()

    def resample(self, particles):
        w = np.array(self.f.weight)
        u01 = float(np.random.random()) if np.max(w) - np.min(w) > 200 else 0.5
        did, anc = self.f.resample(u01)
        self.ancestors.append(anc.copy())
        return self.views
