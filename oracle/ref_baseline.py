"""TEST / BENCH INFRASTRUCTURE -- times the UNMODIFIED Python reference.

bench.py's `cpu_baseline` (kind "reference") and `--impl reference` arm: the
reference's own Robot / HybridMap / GridMap / Scan classes and main.resample, imported
verbatim from oracle/_ref (oracle/vendor_ref.py), driven like main.py:138-166 drives
them -- [p.imu_update(reading) ...], [p.map_update(scan, last_scan, False) ...],
resample(particles) -- on a bounded sample of the bench's synthetic workload.

The one thing the reference cannot bring along is MATLAB: `eng.matchScanCustom`
(hybridmap.py:244-251) is answered by the oracle's restated matcher (C), run on a
shadow oracle map kept in step with the particle's HybridMap.  That stand-in costs
milliseconds against the second or so of pure Python per particle-scan, so the
figure is, if anything, kind to the reference.  Never imported by thesis_b200/.
"""
import contextlib
import io
import multiprocessing as mp
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "robot.py"))


def _install():
    for name in ("matplotlib", "matplotlib.pyplot", "matlab", "matlab.engine"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matlab"].double = lambda x: x               # hybridmap.py:245 wraps lists in matlab.double
    sys.modules["matlab"].engine = sys.modules["matlab.engine"]
    for p in (REF, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


class _Engine:
    """Answers eng.matchScanCustom for ONE robot with the oracle's restated matcher."""

    def __init__(self, O):
        self.O = O
        self.shadow = O.Map()
        self.guess = None
        self.scan = None

    def matchScanCustom(self, curr, ref, guess0, res, prange, nargout=3):
        r = self.shadow.match(self.guess, self.scan, float(prange[0]), float(prange[1]))
        corr = r["pose"] - self.guess
        return [list(corr)], r["cov"].tolist(), r["score"]


def _fresh_robot(O):
    """Robot(eng) with a private tile list: the class-level list at hybridmap.py:64 would alias the maps
    of all particles (a declared deviation that oracle and CUDA path make too)."""
    import hybridmap
    import robot

    eng = _Engine(O)
    hybridmap.HybridMap._maps = []
    r = robot.Robot(eng)
    r._map._maps = list(r._map._maps)
    hybridmap.HybridMap._maps = []
    r._cov = np.zeros((3, 3), dtype=np.float64)              # float64, like oracle and CUDA path
    return r


def _run(args, hook=None):
    ranges, angles, odom, dt, n_particles, n_scans, seed = args
    _install()
    import oracle as O
    import lidar
    import main as refmain
    import models
    from IntelRawIMUData import IntelRawIMUData

    np.random.seed(seed)
    sink = io.StringIO()
    particles = [_fresh_robot(O) for _ in range(n_particles)]
    scan0 = lidar.Scan(np.asarray(ranges[0]), np.asarray(angles), 0)
    oscan0 = O.Scan(ranges[0], angles)
    with contextlib.redirect_stdout(sink):
        for p in particles:                                   # map seeding, main.py:89-90
            for _ in range(2):
                p._map.update(p.get_latest_pose(), scan0)
                p._map._matlab.shadow.update(np.zeros(3), oscan0)
    t0 = time.perf_counter()
    for s in range(1, n_scans + 1):
        if hook is not None:
            hook(s - 1)
        reading = models.Reading(np.asarray(odom[s - 1], dtype=np.float64), s, IntelRawIMUData.progress_pose,
                                 IntelRawIMUData.get_cov_change_matrix, IntelRawIMUData.get_cov_input_uncertainty)
        reading.set_dt(dt * 1e4)                              # main.py:142
        scan = lidar.Scan(np.asarray(ranges[s]), np.asarray(angles), s)
        oscan = O.Scan(ranges[s], angles)
        with contextlib.redirect_stdout(sink):
            [p.imu_update(reading) for p in particles]        # main.py:144
            for p in particles:                               # main.py:157
                eng = p._map._matlab
                eng.guess = np.array([p._x[-1], p._y[-1], p._theta[-1]], dtype=np.float64)
                eng.scan = oscan
                p.map_update(scan, None, False)
                eng.shadow.update(np.array([p._x[-1], p._y[-1], p._theta[-1]], dtype=np.float64), oscan)
                p._cov = np.array(p._cov, dtype=np.float64)
            for p in particles:
                p._weight = [np.float64(w) for w in p._weight]
            particles = refmain.resample(particles)           # main.py:160
        seen = set()
        for p in particles:                                   # a copy shares its original's engine object: give it its own shadow
            eng = p._map._matlab
            if id(eng) in seen:
                e2 = _Engine(O)
                e2.shadow = eng.shadow.copy()
                p._map._matlab = e2
                for m in p._map._maps:
                    m._matlab = e2
            seen.add(id(p._map._matlab))
        sink.seek(0)
        sink.truncate()
    return time.perf_counter() - t0


def _run_steps(args):
    """Worker of run_steps: warm-up scans, barrier, timed scans, barrier."""
    ranges, angles, odom, dt, per, warm, steps, seed, bar = args
    import threading  # noqa: F401  (barrier is a multiprocessing.Barrier)

    t = {}

    def hook(s):
        if s == warm:                                          # all workers start the timed scans together
            bar.wait()
            t["t0"] = time.perf_counter()

    _run((ranges, angles, odom, dt, per, warm + steps, seed), hook)
    t["t1"] = time.perf_counter()
    return t["t1"] - t["t0"]


def run_steps(ranges, angles, odom, dt, particles_per_worker, warmup, steps, workers, seed=11):
    """--impl reference arm: `workers` processes, each a small filter of the unmodified reference,
    `warmup` untimed scans and then `steps` timed scans each.  Returns (seconds, total particles)."""
    if not available():
        return None
    ctx = mp.get_context("fork")
    bar = ctx.Barrier(workers)
    jobs = [(ranges, angles, odom, dt, particles_per_worker, warmup, steps, seed + i, bar) for i in range(workers)]
    procs, q = [], ctx.Queue()

    def target(job):
        q.put(_run_steps(job))

    for j in jobs:
        pr = ctx.Process(target=target, args=(j,))
        pr.start()
        procs.append(pr)
    times = [q.get() for _ in procs]
    for pr in procs:
        pr.join()
    return max(times), particles_per_worker * workers


def run(ranges, angles, odom, dt, n_particles, n_scans, workers=1, seed=11):
    """updates/s of the Python reference: n_particles x n_scans of the workload.  workers > 1 fans
    independent groups of particles over processes (the reference itself is single-threaded,
    main.py:144,157; particles are independent between resamples, so every group resamples on its own)."""
    if not available():
        return None
    per = max(1, n_particles // workers)
    jobs = [(ranges, angles, odom, dt, per, n_scans, seed + i) for i in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _run(jobs[0])
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_run, jobs)
    wall = time.perf_counter() - t0
    return dict(updates_per_s=per * workers * n_scans / wall, particles=per * workers, scans=n_scans, workers=workers, seconds=wall)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    from thesis_b200 import synth

    w = synth.Workload(4, int(sys.argv[2]) if len(sys.argv) > 2 else 360)
    print(run(w.ranges, w.angles, w.odom, w.dt, int(sys.argv[1]) if len(sys.argv) > 1 else 2, 2, workers=int(sys.argv[3]) if len(sys.argv) > 3 else 1))
