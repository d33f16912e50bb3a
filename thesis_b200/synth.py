"""Synthetic workload of BASELINE.json configs[4]:
a 200 m x 200 m office-like world on the reference's 0.05 m lattice, a 360-beam
180-degree lidar ray-cast against it, a closed-loop trajectory sampled at the
reference's update gate (>= 0.33 m or >= pi/9 between scans, main.py:41-43,155)
and velocity-family odometry (IntelRawIMUData.py:33-55 model) with noise.

Everything is generated with numpy from fixed seeds so that the GPU arm, the
CPU-baseline arm and the tests see identical inputs.  Host-side data
preparation only -- nothing here is on the timed path.
"""
import numpy as np

CELL = 0.05
WORLD_M = 200.0
WORLD_CELLS = int(WORLD_M / CELL)            # 4000
MAX_RANGE = 81.83                            # the SICK's "no return" value in the CARMEN logs
VEL_PAR = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)   # IntelRawIMUData.py:48-55


def make_world(seed=0, clutter_fraction=0.0005):
    """Occupancy grid [ix, iy] (True = obstacle), origin at the world centre.

    Axis-aligned walls every 10 m (at +-5, +-15, ... so that the origin is a room
    centre) with a 1 m door in the middle of every 10 m wall segment, plus
    random single-cell clutter."""
    rng = np.random.default_rng(seed)
    g = np.zeros((WORLD_CELLS, WORLD_CELLS), dtype=bool)
    half = WORLD_CELLS // 2
    wall_pos = np.arange(-95.0, 96.0, 10.0)
    coord = (np.arange(WORLD_CELLS) - half + 0.5) * CELL            # cell centres
    # door mask along a wall: 1 m gap centred on multiples of 10 m
    off = np.abs(((coord + 5.0) % 10.0) - 5.0)
    door = off < 0.5
    for w in wall_pos:
        k = int(np.floor(w / CELL)) + half
        g[k, ~door] = True
        g[~door, k] = True
    g[0, :] = g[-1, :] = g[:, 0] = g[:, -1] = True                  # outer boundary
    clutter = rng.random(g.shape) < clutter_fraction
    # keep the door lines and the room-centre lines the robot drives on free of clutter
    centre_line = np.abs(((coord + 5.0) % 10.0) - 5.0) < 0.6
    clutter[centre_line, :] = False
    clutter[:, centre_line] = False
    g |= clutter
    return g


def beam_angles(n_beams):
    """[-pi/2, pi/2] like the CARMEN loaders (FreidLidarData.py:20)."""
    return np.array([-np.pi / 2 + i * np.pi / (n_beams - 1) for i in range(n_beams)])


def raycast(world, pose, angles, max_range=MAX_RANGE, step=CELL / 2):
    """Exact-enough ranges by marching each beam through the ground truth."""
    half = world.shape[0] // 2
    x, y, th = pose
    a = angles + th
    ca, sa = np.cos(a), np.sin(a)
    n = int(max_range / step)
    r = np.full(len(angles), max_range)
    alive = np.ones(len(angles), dtype=bool)
    d = 0.0
    for _ in range(n):
        d += step
        if not alive.any():
            break
        idx = np.flatnonzero(alive)
        ix = np.floor((x + d * ca[idx]) / CELL).astype(np.int64) + half
        iy = np.floor((y + d * sa[idx]) / CELL).astype(np.int64) + half
        inside = (ix >= 0) & (ix < world.shape[0]) & (iy >= 0) & (iy < world.shape[1])
        hit = np.zeros(len(idx), dtype=bool)
        hit[inside] = world[ix[inside], iy[inside]]
        hit |= ~inside
        r[idx[hit]] = d
        alive[idx[hit]] = False
    return r


def trajectory(n_scans, step_m=0.35, loop_m=40.0):
    """Closed rectangular loop through the door centres: (0,0) -> (L,0) -> (L,L)
    -> (0,L) -> (0,0), turning in place in steps of pi/10 at the corners."""
    poses = [(0.0, 0.0, 0.0)]
    corners = [(loop_m, 0.0), (loop_m, loop_m), (0.0, loop_m), (0.0, 0.0)]
    x, y, th = 0.0, 0.0, 0.0
    ci = 0
    while len(poses) < n_scans:
        tx, ty = corners[ci % 4]
        want = np.arctan2(ty - y, tx - x)
        dth = (want - th + np.pi) % (2 * np.pi) - np.pi
        if abs(dth) > 1e-9:
            th += np.clip(dth, -np.pi / 10, np.pi / 10)
        else:
            dist = np.hypot(tx - x, ty - y)
            s = min(step_m, dist)
            x += s * np.cos(th)
            y += s * np.sin(th)
            if dist - s < 1e-9:
                ci += 1
        poses.append((x, y, th))
    return np.array(poses[:n_scans])


class Workload:
    """n_scans sweeps + odometry increments for one run."""

    def __init__(self, n_scans, n_beams=360, seed=0, noise_m=0.01):
        self.world = make_world(seed)
        self.angles = beam_angles(n_beams)
        self.truth = trajectory(n_scans)
        rng = np.random.default_rng(seed + 1)
        self.ranges = np.empty((n_scans, n_beams))
        for i, p in enumerate(self.truth):
            r = raycast(self.world, p, self.angles)
            r = np.where(r < MAX_RANGE, r + rng.normal(0, noise_m, n_beams), r)
            self.ranges[i] = np.clip(r, 0.0, MAX_RANGE)
        # world-frame velocity odometry with dt = 1 s (additive-velocity family)
        d = np.diff(self.truth, axis=0)
        # noise well above the matcher lattice (0.05 m, 0.26 deg) so that the zero correction --
        # which isValidPose rejects (matchScanCustom.m:55) -- is rarely the best match
        self.odom = d + rng.normal(0, [0.03, 0.03, 0.02], d.shape)
        self.dt = 1.0
        self.par = VEL_PAR

    def ray_cells(self, i):
        """A_r of SURVEY 8d for scan i: cells written by the ray-cast."""
        return float(np.sum(np.minimum(self.ranges[i], 15.0) / CELL))


def write_carmen_log(path, work, t0=100.0, dt=0.5, odom_per_scan=4):
    """Write the workload as a CARMEN log (ODOM / FLASER lines) so that the
    reference-style loaders (thesis_b200.loaders) and the headless main.py loop can
    be exercised on logs that are not in the reference tree (aces.txt, fr.log, ...
    are missing there, SURVEY 8c "data gaps").  Odometry = ground truth + the
    workload's noise, interpolated odom_per_scan times between sweeps."""
    odo = np.vstack(([0.0, 0.0, 0.0], np.cumsum(work.odom, axis=0)))
    lines = []
    for i in range(len(work.ranges)):
        ts = t0 + i * dt
        if i > 0:
            for k in range(1, odom_per_scan + 1):
                a = k / odom_per_scan
                p = odo[i - 1] * (1 - a) + odo[i] * a
                tk = t0 + (i - 1) * dt + a * dt - 1e-3
                lines.append("ODOM %.6f %.6f %.6f 0 0 0 %.6f synth %.6f" % (p[0], p[1], p[2], tk, tk))
        else:
            lines.append("ODOM 0 0 0 0 0 0 %.6f synth %.6f" % (ts - 1e-3, ts - 1e-3))
        r = " ".join("%.3f" % v for v in work.ranges[i])
        p = odo[i]
        lines.append("FLASER %d %s %.6f %.6f %.6f %.6f %.6f %.6f %.6f synth %.6f" % (
            work.ranges.shape[1], r, p[0], p[1], p[2], p[0], p[1], p[2], ts, ts))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return path
