"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/rbpf_oracle.c.

The oracle is the CPU restatement of the reference's per-scan RBPF update (see
the header of rbpf_oracle.c for what is pinned and what is not).  Importers:
tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
Nothing under thesis_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DIM = 800
TILE_LEN = 40
CS = 0.05
MAX_NT = 14
SLICE_W = 2 * MAX_NT + 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "rbpf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_map_new.restype = C.c_void_p
        L.orc_map_copy.restype = C.c_void_p
        L.orc_map_copy.argtypes = [C.c_void_p]
        L.orc_map_free.argtypes = [C.c_void_p]
        L.orc_map_ntiles.argtypes = [C.c_void_p]
        L.orc_map_tile_centre.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        L.orc_map_tile_cells.restype = _dp
        L.orc_map_tile_cells.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_odds_at.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp]
        L.orc_scan_prepare.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_transform.argtypes = [_dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_bresenham.argtypes = [C.c_int] * 4 + [_ip, C.c_int]
        L.orc_map_update.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_nearby_occ.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, C.c_int]
        L.orc_match_ref.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int]
        L.orc_sample_weight.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_propose.argtypes = [_dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_moments.restype = C.c_double
        L.orc_moments.argtypes = [_dp, _dp, C.c_int, _dp, _dp]
        L.orc_rot_step.restype = C.c_double
        L.orc_window_cells.argtypes = [C.c_double]
        L.orc_pose_range.argtypes = [_dp, _dp, _dp]
        L.orc_match.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                _dp, _dp, _dp, _ip, _ip]
        L.orc_match_adj.argtypes = [_dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                    _dp, _dp, _dp, _ip, _ip]
        L.orc_filter_map_update_adj.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.orc_match_curr.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, _dp]
        L.orc_ndt_probe.argtypes = [_dp, _dp]
        L.orc_ndt_probe.restype = None
        L.orc_motion.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp, _dp]
        L.orc_resample.argtypes = [_dp, C.c_int, C.c_double, _ip]
        L.orc_filter_new.restype = C.c_void_p
        L.orc_filter_new.argtypes = [C.c_int] * 3
        L.orc_filter_free.argtypes = [C.c_void_p]
        for n, t in (("pose", _dp), ("cov", _dp), ("weight", _dp), ("valid", _ip)):
            f = getattr(L, "orc_filter_" + n)
            f.restype = t
            f.argtypes = [C.c_void_p]
        L.orc_filter_map.restype = C.c_void_p
        L.orc_filter_map.argtypes = [C.c_void_p, C.c_int]
        L.orc_filter_set_scan.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_filter_motion.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, _dp]
        L.orc_filter_integrate.argtypes = [C.c_void_p]
        L.orc_rb_sincos.argtypes = [_dp, C.c_int, _dp, _dp]
        L.orc_rb_exp.argtypes = [_dp, C.c_int, _dp]
        L.orc_filter_map_update.argtypes = [C.c_void_p, _dp]
        L.orc_filter_map_update_guesses.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.orc_propose_pdf.argtypes = [_dp, _dp, _dp, C.c_int, _dp]
        L.orc_filter_resample.argtypes = [C.c_void_p, C.c_double, _ip]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Scan:
    """Beam geometry of one sweep (lidar.py:76-80): px, py, dist."""

    def __init__(self, ranges, angles):
        r, rp = _d(ranges)
        a, ap = _d(angles)
        self.B = len(r)
        self.px = np.empty(self.B)
        self.py = np.empty(self.B)
        self.dist = np.empty(self.B)
        lib().orc_scan_prepare(rp, ap, self.B, _d(self.px)[1], _d(self.py)[1], _d(self.dist)[1])

    def ptrs(self):
        return _d(self.px)[1], _d(self.py)[1], _d(self.dist)[1]


class Map:
    """One particle's tile list (hybridmap.py:63-70)."""

    def __init__(self, handle=None, owned=True):
        self._h = C.c_void_p(lib().orc_map_new() if handle is None else handle)
        self._owned = owned

    def __del__(self):
        if getattr(self, "_owned", False) and self._h:
            lib().orc_map_free(self._h)
            self._h = None

    def copy(self):
        return Map(lib().orc_map_copy(self._h))

    def tiles(self):
        """{(cx, cy): 800x800 float64 view [ix][iy]} in allocation order."""
        out = {}
        for i in range(lib().orc_map_ntiles(self._h)):
            cx, cy = C.c_int(), C.c_int()
            lib().orc_map_tile_centre(self._h, i, C.byref(cx), C.byref(cy))
            p = lib().orc_map_tile_cells(self._h, i)
            out[(cx.value, cy.value)] = np.ctypeslib.as_array(p, shape=(DIM, DIM))
        return out

    def odds_at(self, x, y):
        v = C.c_double()
        return v.value if lib().orc_get_odds_at(self._h, x, y, C.byref(v)) else None

    def update(self, pose, scan):
        p, pp = _d(pose)
        lib().orc_map_update(self._h, pp, *scan.ptrs(), scan.B)

    def match_ref(self, guess, scan, cap=40000):
        """valid_ref_points of hybridmap.py:230-239 (relative to the guess) as the restated matcher masks them."""
        out = np.empty((cap, 2))
        n = lib().orc_match_ref(self._h, _d(guess)[1], _d(scan.px)[1], _d(scan.py)[1], _d(scan.dist)[1], len(scan.px), _d(out)[1], cap)
        return out[:n]

    def nearby_occ(self, x, y, cap=20000):
        out = np.empty((cap, 2))
        n = lib().orc_nearby_occ(self._h, x, y, _d(out)[1], cap)
        assert n <= cap
        return out[:n].copy()

    def sample_weight(self, guesses, scan, prs):
        g, gp = _d(guesses)
        p, ppr = _d(prs)
        K = len(p)
        w = np.empty(K)
        lib().orc_sample_weight(self._h, gp, K, *scan.ptrs(), scan.B, ppr, _d(w)[1])
        return w

    def match_curr(self, guess, scan):
        out = np.empty((scan.B, 2))
        M = lib().orc_match_curr(self._h, _d(guess)[1], *scan.ptrs(), scan.B, _d(out)[1])
        return out[:M].copy()

    def match(self, guess, scan, rx, ry):
        g, gp = _d(guess)
        pose = np.empty(3)
        cov = np.empty(9)
        score = C.c_double()
        dbg = np.zeros(8, dtype=np.int32)
        sl = np.zeros(SLICE_W * SLICE_W, dtype=np.int32)
        valid = lib().orc_match(self._h, gp, *scan.ptrs(), scan.B, rx, ry, _d(pose)[1], _d(cov)[1],
                                C.byref(score), _i(dbg), _i(sl))
        return dict(valid=bool(valid), pose=pose, cov=cov.reshape(3, 3), score=score.value,
                    M=int(dbg[0]), best=(int(dbg[1]), int(dbg[2]), int(dbg[3])), nx=int(dbg[4]),
                    ny=int(dbg[5]), ndt_evals=int(dbg[6]), ndt_accepted=bool(dbg[7]), slice=sl.reshape(SLICE_W, SLICE_W))


def match_adj(guess, scan, prev_xy, rx, ry):
    """Scan-to-previous-scan matcher (hybridmap.py:147-191); prev_xy = [n, 2] global endpoints."""
    prev = np.ascontiguousarray(prev_xy, dtype=np.float64)
    px, py = np.ascontiguousarray(prev[:, 0]), np.ascontiguousarray(prev[:, 1])
    pose = np.empty(3)
    cov = np.empty(9)
    score = C.c_double()
    dbg = np.zeros(8, dtype=np.int32)
    sl = np.zeros(SLICE_W * SLICE_W, dtype=np.int32)
    valid = lib().orc_match_adj(_d(guess)[1], _d(scan.px)[1], _d(scan.py)[1], scan.B, _d(px)[1], _d(py)[1], len(px),
                                rx, ry, _d(pose)[1], _d(cov)[1], C.byref(score), _i(dbg), _i(sl))
    return dict(valid=bool(valid), pose=pose, cov=cov.reshape(3, 3), score=score.value, M=int(dbg[0]),
                best=(int(dbg[1]), int(dbg[2]), int(dbg[3])), nx=int(dbg[4]), ny=int(dbg[5]),
                ndt_evals=int(dbg[6]), ndt_accepted=bool(dbg[7]), slice=sl.reshape(SLICE_W, SLICE_W))


def set_refine(on):
    """Switch the NDT stage (matchScanCustom.m:32-50) of match()/match_adj()/Filter on or off; returns the old value."""
    return bool(lib().orc_set_refine(int(bool(on))))


def ndt_terms(m, guess, scan, p):
    """{S, gradient, Hessian, curvature model} of the NDT score of map m at correction p = (cells, cells, rad)."""
    out = np.zeros(16)
    pp = np.ascontiguousarray(p, dtype=np.float64)
    lib().orc_ndt_probe(_d(pp)[1], _d(out)[1])
    try:
        m.match(guess, scan, 0.7, 0.7)
    finally:
        lib().orc_ndt_probe(None, None)
    H = np.array([[out[4], out[5], out[6]], [out[5], out[7], out[8]], [out[6], out[8], out[9]]])
    Cm = np.array([[out[10], out[11], out[12]], [out[11], out[13], out[14]], [out[12], out[14], out[15]]])
    return dict(S=out[0], grad=out[1:4].copy(), hess=H, model=Cm)


def set_threads(n=0):
    """OpenMP threads used by Filter (n <= 0: query only)."""
    return lib().orc_set_threads(int(n))


def transform(pose, scan):
    gx = np.empty(scan.B)
    gy = np.empty(scan.B)
    lib().orc_transform(_d(pose)[1], _d(scan.px)[1], _d(scan.py)[1], scan.B, _d(gx)[1], _d(gy)[1])
    return gx, gy


def bresenham(x0, y0, x1, y1):
    cap = 2 * (abs(x1 - x0) + abs(y1 - y0)) + 8
    out = np.empty((cap, 2), dtype=np.int32)
    n = lib().orc_bresenham(x0, y0, x1, y1, _i(out), cap)
    return out[:n].copy()


def propose(mean, cov, z):
    z, zp = _d(z)
    K = z.shape[0]
    g = np.empty((K, 3))
    prs = np.empty(K)
    lib().orc_propose(_d(mean)[1], _d(np.asarray(cov).reshape(9))[1], zp, K, _d(g)[1], _d(prs)[1])
    return g, prs


def rb_sincos(a):
    """csrc/rb_math.h rb_sincos over an array (the routine the CUDA kernels and this oracle share)."""
    a, ap = _d(a)
    s, c = np.empty_like(a), np.empty_like(a)
    lib().orc_rb_sincos(ap, a.size, _d(s)[1], _d(c)[1])
    return s, c


def rb_exp(x):
    x, xp = _d(x)
    out = np.empty_like(x)
    lib().orc_rb_exp(xp, x.size, _d(out)[1])
    return out


def propose_numpy(mean, cov, K=30):
    """robot.py:81 with NumPy itself -- `np.random.multivariate_normal(scan_pose, scan_cov, 30)` on the
    global legacy stream -- and robot.py:87's densities for those samples (orc_propose_pdf)."""
    g = np.ascontiguousarray(np.random.multivariate_normal(np.asarray(mean, dtype=np.float64), np.asarray(cov, dtype=np.float64), K))
    prs = np.empty(K)
    lib().orc_propose_pdf(_d(mean)[1], _d(np.asarray(cov).reshape(9))[1], _d(g)[1], K, _d(prs)[1])
    return g, prs


def moments(guesses, w):
    g, gp = _d(guesses)
    w, wp = _d(w)
    mean = np.empty(3)
    sigma = np.empty(9)
    norm = lib().orc_moments(gp, wp, len(w), _d(mean)[1], _d(sigma)[1])
    return mean, sigma.reshape(3, 3), norm


def pose_range(cov):
    rx, ry = C.c_double(), C.c_double()
    lib().orc_pose_range(_d(np.asarray(cov).reshape(9))[1], C.byref(rx), C.byref(ry))
    return rx.value, ry.value


def rot_step():
    return lib().orc_rot_step()


def rot_count():
    return lib().orc_rot_count()


def motion(family, u, dt, par, pose, cov):
    pose = np.array(pose, dtype=np.float64)
    cov = np.array(cov, dtype=np.float64).reshape(9)
    u4 = np.zeros(4)
    u4[: len(u)] = u
    p4 = np.zeros(4)
    p4[: len(par)] = par
    lib().orc_motion(family, _d(u4)[1], dt, _d(p4)[1], _d(pose)[1], _d(cov)[1])
    return pose, cov.reshape(3, 3)


def resample(weights, u01):
    """(status, ancestors): 0 no resample, 1 resampled, -1 reference AssertionError."""
    w, wp = _d(weights)
    anc = np.empty(len(w), dtype=np.int32)
    rc = lib().orc_resample(wp, len(w), u01, _i(anc))
    return rc, anc


class Filter:
    """N particles driven like main.py:138-166 (CPU, OpenMP over particles)."""

    def __init__(self, N, B, K=30):
        self.N, self.B, self.K = N, B, K
        self._h = C.c_void_p(lib().orc_filter_new(N, B, K))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_filter_free(self._h)
            self._h = None

    def _arr(self, name, shape, dtype=np.float64):
        return np.ctypeslib.as_array(getattr(lib(), "orc_filter_" + name)(self._h), shape=shape)

    @property
    def pose(self):
        return self._arr("pose", (self.N, 3))

    @property
    def cov(self):
        return self._arr("cov", (self.N, 3, 3))

    @property
    def weight(self):
        return self._arr("weight", (self.N,))

    @property
    def valid(self):
        return self._arr("valid", (self.N,))

    def map(self, i):
        return Map(lib().orc_filter_map(self._h, i), owned=False)

    def set_scan(self, ranges, angles):
        lib().orc_filter_set_scan(self._h, _d(ranges)[1], _d(angles)[1])

    def motion(self, family, u, dt, par=(0, 0, 0, 0)):
        u4 = np.zeros(4)
        u4[: len(u)] = u
        p4 = np.zeros(4)
        p4[: len(par)] = par
        lib().orc_filter_motion(self._h, family, _d(u4)[1], dt, _d(p4)[1])

    def integrate(self):
        lib().orc_filter_integrate(self._h)

    def map_update(self, z, prev_xy=None, guesses=False):
        """z: N*K*3 standard normals for the mean + chol(cov) z transform, or -- guesses=True -- the
        proposal samples themselves, drawn by the caller like robot.py:81 (see propose_numpy)."""
        z, zp = _d(z)
        assert z.size == self.N * self.K * 3
        if guesses:
            if prev_xy is None:
                lib().orc_filter_map_update_guesses(self._h, zp, None, None, 0)
            else:
                prev = np.ascontiguousarray(prev_xy, dtype=np.float64)
                px, py = np.ascontiguousarray(prev[:, 0]), np.ascontiguousarray(prev[:, 1])
                lib().orc_filter_map_update_guesses(self._h, zp, _d(px)[1], _d(py)[1], len(px))
        elif prev_xy is None:
            lib().orc_filter_map_update(self._h, zp)
        else:
            prev = np.ascontiguousarray(prev_xy, dtype=np.float64)
            px, py = np.ascontiguousarray(prev[:, 0]), np.ascontiguousarray(prev[:, 1])
            lib().orc_filter_map_update_adj(self._h, zp, _d(px)[1], _d(py)[1], len(px))

    def resample(self, u01):
        anc = np.empty(self.N, dtype=np.int32)
        rc = lib().orc_filter_resample(self._h, u01, _i(anc))
        if rc < 0:
            raise AssertionError("Incorrect number of resampled weights.")
        return bool(rc), anc
