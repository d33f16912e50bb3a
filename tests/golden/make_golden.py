"""Generate tests/golden/ref_golden.npz by RUNNING THE PYTHON REFERENCE.

Only works where /root/reference exists (the build container); the resulting
.npz is committed so that the CPU tests can pin oracle/rbpf_oracle.c -- and
through it the CUDA path -- on the GPU box, where the reference is absent.

Every array is produced by the reference's own functions imported through
oracle/ref_shim.py (matplotlib / matlab stubbed).  Usage:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.install()
with ref_shim.ref_cwd():
    from IntelLidarData import IntelLidarData
    from IntelIMUData import IntelIMUData
    from lidar import Lidar
    from imu import IMU

    LD = Lidar(IntelLidarData(), None)
    IMU_INTEL = IMU(IntelIMUData())

import hybridmap  # noqa: E402
import main as refmain  # noqa: E402
import models  # noqa: E402
from DefaultIMUData import DefaultIMUData  # noqa: E402
from IntelRawIMUData import IntelRawIMUData  # noqa: E402
from AcesIMUData import AcesIMUData  # noqa: E402

N_SCANS = 40
G = {}
rng = np.random.default_rng(20261018)

# -- the Intel excerpt every other vector is built on (IntelLidarData.py:12-20)
G["intel_ranges"] = LD._scans[:N_SCANS].copy()
G["intel_angles"] = LD._angles.copy()

# -- beam geometry, lidar.py:76-80 and :111-128
G["scan0_xy"] = np.stack([LD[0].x(), LD[0].y()])
xf_pose = np.array([1.2345, -0.777, 0.4321])
g = LD[3].from_global_reference(models.Pose(*xf_pose))
G["xform_pose"] = xf_pose
G["xform_xy"] = np.stack([g.x(), g.y()])

# -- Bresenham incl. degenerate rays, hybridmap.py:274-301
rays = rng.integers(-60, 60, size=(400, 4))
rays[:8] = [(5, 5, 5, 2), (5, 5, 2, 5), (3, 3, 3, 3), (5, 5, 5, 9), (5, 5, 9, 5), (0, 0, 7, 7), (0, 0, -7, 7), (2, 1, -9, -4)]
cells, offs = [], [0]
for x0, y0, x1, y1 in rays:
    c = hybridmap.HybridMap.get_affected_points(int(x0), int(y0), int(x1), int(y1))
    cells.extend(c)
    offs.append(len(cells))
G["bres_rays"] = rays.astype(np.int32)
G["bres_cells"] = np.array(cells, dtype=np.int32).reshape(-1, 2)
G["bres_offs"] = np.array(offs, dtype=np.int64)


def sparse_tiles(hmap, prefix):
    """Store every tile as (centre, flat nonzero index, value)."""
    cen = []
    for n, m in enumerate(hmap._maps):
        a = m.map()._map
        idx = np.flatnonzero(a)
        cen.append((m.centre().x, m.centre().y))
        G["%s_t%d_idx" % (prefix, n)] = idx.astype(np.int32)
        G["%s_t%d_val" % (prefix, n)] = a.ravel()[idx].copy()
    G[prefix + "_centres"] = np.array(cen, dtype=np.int32)


# -- map integration across tile borders and negative coordinates, hybridmap.py:95-145
traj = np.array([(0, 0, 0), (0, 0, 0), (-3.3, 2.1, 1.0), (-18.7, -17.9, 2.5), (19.2, -19.8, -2.0),
                 (-22.0, 5.0, 3.0), (35.5, 21.0, 0.3)], dtype=np.float64)
traj_scan = np.array([0, 0, 3, 6, 9, 12, 15], dtype=np.int32)
hm = ref_shim.fresh_hybridmap()
for p, si in zip(traj, traj_scan):
    hm.update(models.Pose(*p), LD[int(si)])
G["integ_poses"] = traj
G["integ_scan_idx"] = traj_scan
sparse_tiles(hm, "integ")

# -- a long-range synthetic scan exercising the 15 m clip (:107-113) and r ~ 0
long_ranges = rng.uniform(0.0, 40.0, 180)
long_ranges[:4] = [0.0, 1e-4, 15.0, 15.0000001]
hm2 = ref_shim.fresh_hybridmap()
import lidar as reflidar  # noqa: E402

ls = reflidar.Scan(long_ranges, LD._angles, 0)
clip_pose = np.array([-7.31, 11.9, -0.8])
hm2.update(models.Pose(*clip_pose), ls)
G["clip_ranges"] = long_ranges
G["clip_pose"] = clip_pose
sparse_tiles(hm2, "clip")

# -- read lookups, hybridmap.py:85-93 (None -> NaN)
pts = rng.uniform(-45, 45, (3000, 2))
pts[:200] = np.round(pts[:200] / 0.05) * 0.05          # lattice-aligned probes
vals = []
for x, y in pts:
    v = hm.get_odds_at(models.Position(x, y))
    vals.append(np.nan if v is None else v)
G["odds_pts"] = pts
G["odds_vals"] = np.array(vals)

# -- one particle: map seeded by scans 0,0,1,2 ; then a full Robot.map_update (robot.py:59-115)
r = ref_shim.fresh_robot()
seed_poses = np.array([(0, 0, 0), (0, 0, 0), (0.3, 0.1, 0.05), (0.6, 0.15, 0.1)], dtype=np.float64)
for i, p in enumerate(seed_poses):
    r._map.update(models.Pose(*p), LD[[0, 0, 1, 2][i]])
G["upd_seed_poses"] = seed_poses
G["upd_seed_scan_idx"] = np.array([0, 0, 1, 2], dtype=np.int32)

# sample weights, robot.py:118-139
guesses = np.array([0.6, 0.15, 0.1]) + rng.normal(0, [0.02, 0.02, 0.005], (30, 3))
prs = rng.uniform(1, 1e6, 30)
G["sw_guesses"] = guesses
G["sw_prs"] = prs
G["sw_scan_idx"] = np.int32(4)
G["sw_w"] = np.array(r._generate_sample_weight(guesses, LD[4], prs), dtype=np.float64)


class FakeEngine:
    """Stands in for the MATLAB engine (hybridmap.py:244-251): records the point
    sets the reference builds and returns a canned matcher answer."""

    def __init__(self, pose, cov, score):
        self.answer = (pose, cov, score)

    def matchScanCustom(self, curr, ref, guess, res, prange, nargout=3):
        self.curr = np.array(curr, dtype=np.float64)
        self.ref = np.array(ref, dtype=np.float64).reshape(-1, 2)
        self.res = res
        self.prange = np.array(prange, dtype=np.float64)
        return [list(self.answer[0])], self.answer[1], self.answer[2]


m_corr = np.array([0.02, -0.03, 0.01])
m_cov = np.diag([4e-4, 3e-4, 2e-5])
m_cov[0, 1] = m_cov[1, 0] = 1e-4
eng = FakeEngine(m_corr, m_cov.tolist(), 55.0)
r._map._matlab = eng
r._x.append(0.6)
r._y.append(0.15)
r._theta.append(0.1)
r._cov = np.diag([1e-4, 1e-4, 1e-6])
G["upd_prior_cov"] = np.array(r._cov)
G["upd_match_corr"] = m_corr
G["upd_match_cov"] = m_cov
# The reference draws its proposal samples with np.random.multivariate_normal on the global legacy stream
# (robot.py:81).  Nothing is patched: the stream is seeded, Robot.map_update runs as it is, and the samples
# it drew are reproduced afterwards by seeding again and making the same call.
UPD_SEED = 20261018
G["upd_seed"] = np.int64(UPD_SEED)
np.random.seed(UPD_SEED)
r.map_update(LD[4], None, False)
upd_mean = [m_corr[0] + 0.6, m_corr[1] + 0.15, m_corr[2] + 0.1]          # hybridmap.py:253-255
np.random.seed(UPD_SEED)
G["upd_guesses"] = np.array(np.random.multivariate_normal(upd_mean, np.array(m_cov.tolist()), 30))
G["upd_curr"] = eng.curr                     # valid_curr_points, hybridmap.py:240
G["upd_ref"] = eng.ref                       # valid_ref_points,  hybridmap.py:239
G["upd_prange"] = eng.prange
G["upd_pose"] = np.array([r._x[-1], r._y[-1], r._theta[-1]], dtype=np.float64)
G["upd_cov"] = np.array(r._cov, dtype=np.float64)
G["upd_weight"] = np.float64(r._weight[-1])
sparse_tiles(r._map, "upd")

# NaN-covariance fallback, robot.py:73-78
r2 = ref_shim.fresh_robot()
for i, p in enumerate(seed_poses):
    r2._map.update(models.Pose(*p), LD[[0, 0, 1, 2][i]])
r2._map._matlab = FakeEngine(np.zeros(3), (np.full((3, 3), np.nan)).tolist(), 0.0)
r2._x.append(0.6)
r2._y.append(0.15)
r2._theta.append(0.1)
r2._cov = np.diag([1e-4, 1e-4, 1e-6])
r2.map_update(LD[4], None, False)
G["bad_pose"] = np.array([r2._x[-1], r2._y[-1], r2._theta[-1]], dtype=np.float64)
G["bad_weight"] = np.float64(r2._weight[-1])
G["bad_nhist"] = np.int32(len(r2._x))
sparse_tiles(r2._map, "bad")

# -- search window, robot.py:62-65
covs = [np.zeros((3, 3)), np.diag([1e-6, 4e-5, 1]), np.diag([1.0, 1e-7, 1]), np.diag([2.5e-5, 1.2e-5, 0])]
pr = []
for c in covs:
    p = np.sqrt(np.diag(c)) * 30.0
    pr.append([max(min(4 * p[0], 0.7), 0.1), max(min(4 * p[1], 0.7), 0.1)])
G["prange_covs"] = np.array(covs)
G["prange_out"] = np.array(pr)

# -- resampling, main.py:46-79 with float64 weights and a fixed uniform
import io  # noqa: E402
import contextlib  # noqa: E402


class W:
    def __init__(self, w):
        self._weight = [w]
        self.tag = None

    def weight(self):
        return self._weight

    def copy(self):
        c = W(self._weight[-1])
        c.tag = self.tag
        return c


cases = []
for n, scale, shift in ((8, 50.0, 0.0), (64, 1e3, -500.0), (257, 1e6, -9e5), (1000, 1e8, -1e8), (16, 10.0, 0.0), (33, 1e4, 1e4)):
    w = rng.normal(0, 1, n) * scale + shift
    if n == 64:
        w[5] = -np.inf
    seed = int(rng.integers(1, 2 ** 31))
    u = float(np.random.RandomState(seed).random_sample())          # what main.py:59 will draw
    ps = [W(np.float64(x)) for x in w]
    for i, p in enumerate(ps):
        p.tag = i
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        out = refmain.resample(ps)
    anc = np.array([p.tag for p in out], dtype=np.int32)
    did = len(out[0]._weight) == 2
    cases.append((w, u, anc, did))
G["rs_n"] = np.int32(len(cases))
for i, (w, u, anc, did) in enumerate(cases):
    G["rs%d_w" % i] = w
    G["rs%d_u" % i] = np.float64(u)
    G["rs%d_anc" % i] = anc
    G["rs%d_did" % i] = np.bool_(did)

# -- motion families through Robot.imu_update, robot.py:45-57
import robot as refrobot  # noqa: E402


def run_motion(cls, data, dt_ticks, pose0, cov0):
    rb = refrobot.Robot(None)
    rb._x, rb._y, rb._theta = [pose0[0]], [pose0[1]], [pose0[2]]
    rb._cov = np.array(cov0, dtype=np.float64)
    rd = models.Reading(np.array(data, dtype=np.float64), 0, cls.progress_pose, cls.get_cov_change_matrix,
                        cls.get_cov_input_uncertainty)
    rd.set_dt(dt_ticks)
    with contextlib.redirect_stdout(io.StringIO()):
        rb.imu_update(rd)
    return np.array([rb._x[-1], rb._y[-1], rb._theta[-1]], dtype=np.float64), np.array(rb._cov, dtype=np.float64)


cov0 = np.array([[2e-3, 1e-4, -2e-5], [1e-4, 3e-3, 4e-5], [-2e-5, 4e-5, 5e-4]])
pose0 = np.array([1.5, -2.25, 0.7])
mot = [
    ("abs", IntelIMUData, [1.61, -2.2, 0.74], 1000),
    ("velraw", IntelRawIMUData, [0.31, -0.12, 0.2], 1230),
    ("velaces", AcesIMUData, [-0.4, 0.22, -0.31], 870),
    ("uni", DefaultIMUData, [0.83, -0.17], 50),
]
for name, cls, data, dtt in mot:
    p1, c1 = run_motion(cls, data, dtt, pose0, cov0)
    G["mot_%s_u" % name] = np.array(data, dtype=np.float64)
    G["mot_%s_dt_ticks" % name] = np.float64(dtt)
    G["mot_%s_pose" % name] = p1
    G["mot_%s_cov" % name] = c1
G["mot_pose0"] = pose0
G["mot_cov0"] = cov0

# -- end to end: the reference's own Robot / resample / HybridMap classes driven by the
#    headless main.py loop (thesis_b200.harness) over the Intel excerpt, with the restated
#    matcher plugged into the MATLAB seam (tests/ref_adapter.py), np.random.seed(0)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_adapter as RA  # noqa: E402
from thesis_b200 import harness, sensors  # noqa: E402
from excerpt import ExcerptIMU, ExcerptLidar  # noqa: E402

E2E_N, E2E_FRAMES = 4, 12
np.random.seed(0)
ld = sensors.Lidar(ExcerptLidar())
im = sensors.IMU(ExcerptIMU())
rp = [RA.RefParticle(r) for r in RA.make_ref_particles(E2E_N)]
anc_log = []
pose_log = []


def on_frame(frame, parts):
    pose_log.append([[float(v) for v in (p.get_latest_pose().x(), p.get_latest_pose().y(), p.get_latest_pose().theta())]
                     for p in parts])


parts, log = harness.run_log(rp, ld, im, lambda p: RA.ref_resample(p, anc_log), seed_fn=RA.ref_seed,
                             max_frames=E2E_FRAMES, on_frame=on_frame)
G["e2e_n"] = np.int32(E2E_N)
G["e2e_frames"] = np.int32(E2E_FRAMES)
G["e2e_poses"] = np.array(pose_log, dtype=np.float64)            # [frame, particle, 3]
G["e2e_ancestors"] = np.array(anc_log, dtype=np.int32)           # [update, particle]
G["e2e_weights"] = np.array([float(p.weight()[-1]) for p in parts], dtype=np.float64)
G["e2e_updated"] = np.array([l["updated"] for l in log])
for i, p in enumerate(parts[:2]):
    sparse_tiles(p.r._map, "e2e_p%d" % i)

out = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out), "bytes,", len(G), "arrays")
