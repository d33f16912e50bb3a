"""Device-resident particle set and the reference's Robot / resample API on top.

`ParticleSet` is a thin object wrapper over the C ABI (include/rbpf_b200.h); all
arithmetic runs in the CUDA kernels of thesis_b200/csrc.  `Robot` and
`resample` keep the reference's Python signatures (robot.py:19-157,
main.py:46-79) so that a loop written like main.py:138-166 runs unchanged:
particles are *views* (an index into the set).  Because main.py fans
`p.map_update(...)` over the particle list, the first view called in a round
runs the batched kernels for all particles and the remaining calls of that round
are no-ops.
"""
import ctypes as C

import numpy as np

from . import _lib
from .models import Pose

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

DIM = 800
SUBTILE_BYTES = 160 * 160


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class RbpfError(RuntimeError):
    pass


class ParticleSet:
    """N particles on one GPU: poses, covariances, weights, page tables, tile pool."""

    def __init__(self, n_particles, n_beams, n_samples=30, world_tiles=(5, 5), pool_subtiles=None,
                 device=0, rank=0, world=1, stream=0, seed=0):
        self._lib = _lib.load()
        self.N, self.B, self.K = int(n_particles), int(n_beams), int(n_samples)
        self.rank, self.world = rank, world
        if pool_subtiles is None:
            pool_subtiles = max(4096, 48 * self.N)
        cfg = _lib.RbpfConfig(self.N, self.B, self.K, world_tiles[0], world_tiles[1], int(pool_subtiles),
                              device, rank, world, 0, int(stream), int(seed))
        self._h = C.c_void_p()
        rc = self._lib.rbpf_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise RbpfError("rbpf_create failed with status %d (no CUDA device? there is no CPU fallback)" % rc)

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rbpf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.rbpf_last_error(self._h).decode()
            if rc == _lib.RBPF_ERR_RESAMPLE:
                raise AssertionError(msg)                     # main.py:66-67
            raise RbpfError("status %d: %s" % (rc, msg))

    def synchronize(self):
        self._ck(self._lib.rbpf_synchronize(self._h))

    # -- stages (include/rbpf_b200.h)
    def set_scan(self, ranges, angles):
        r, rp = _d(ranges)
        a, ap = _d(angles)
        self._ck(self._lib.rbpf_set_scan(self._h, rp, ap, len(r)))

    def motion(self, family, u, dt, par=(0.0, 0.0, 0.0, 0.0)):
        u4 = np.zeros(4)
        u4[: len(u)] = u
        p4 = np.zeros(4)
        p4[: len(par)] = par
        self._ck(self._lib.rbpf_motion(self._h, int(family), _d(u4)[1], float(dt), _d(p4)[1]))

    def scan_match(self):
        self._ck(self._lib.rbpf_scan_match(self._h))

    def weight(self, z=None):
        if z is None:
            self._ck(self._lib.rbpf_weight(self._h, None))
        else:
            z, zp = _d(z)
            if z.size != self.N * self.K * 3:
                raise ValueError("z must hold N*K*3 standard normals")
            self._ck(self._lib.rbpf_weight(self._h, zp))

    def integrate(self, fallback_weights=False):
        self._ck(self._lib.rbpf_integrate(self._h, 1 if fallback_weights else 0))

    def resample(self, u01=None, want_ancestors=True):
        """Returns (did_resample, ancestors or None)."""
        did = C.c_int32(0)
        anc = np.empty(self.N * self.world, dtype=np.int32) if want_ancestors else None
        up = None
        if u01 is not None:
            u = C.c_double(float(u01))
            up = C.byref(u)
        self._ck(self._lib.rbpf_resample(self._h, C.cast(up, _dp) if up is not None else None,
                                         anc.ctypes.data_as(_ip) if anc is not None else None, C.byref(did)))
        return bool(did.value), anc

    def step(self, ranges, angles):
        """One lidar event, device-side draws, no host synchronisation."""
        r, rp = _d(ranges)
        a, ap = _d(angles)
        self._ck(self._lib.rbpf_step(self._h, rp, ap, len(r)))

    STAGES = ("set_scan", "match", "weight", "raycast_prepare", "raycast_cast", "weight_fallback",
              "resample_plan", "resample_apply")

    def timing_enable(self, max_steps):
        self._ck(self._lib.rbpf_timing_enable(self._h, int(max_steps)))

    def timing_read(self):
        """({stage: total ms}, steps) accumulated by step() since the last read."""
        ms = np.zeros(8)
        n = C.c_int32(0)
        self._ck(self._lib.rbpf_timing_read(self._h, ms.ctypes.data_as(_dp), C.byref(n)))
        return dict(zip(self.STAGES, ms.tolist())), n.value

    # -- state
    def _get(self, fn, shape):
        out = np.empty(shape, dtype=np.float64)
        self._ck(fn(self._h, out.ctypes.data_as(_dp)))
        return out

    @property
    def poses(self):
        return self._get(self._lib.rbpf_get_poses, (self.N, 3))

    @poses.setter
    def poses(self, v):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.N, 3)))
        self._ck(self._lib.rbpf_set_poses(self._h, v.ctypes.data_as(_dp)))

    @property
    def covs(self):
        return self._get(self._lib.rbpf_get_covs, (self.N, 3, 3))

    @covs.setter
    def covs(self, v):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.N, 3, 3)))
        self._ck(self._lib.rbpf_set_covs(self._h, v.ctypes.data_as(_dp)))

    @property
    def weights(self):
        return self._get(self._lib.rbpf_get_weights, (self.N,))

    @weights.setter
    def weights(self, v):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.N,)))
        self._ck(self._lib.rbpf_set_weights(self._h, v.ctypes.data_as(_dp)))

    def match_result(self):
        pose = np.empty((self.N, 3))
        cov = np.empty((self.N, 3, 3))
        score = np.empty(self.N)
        valid = np.empty(self.N, dtype=np.int32)
        best = np.empty((self.N, 4), dtype=np.int32)
        self._ck(self._lib.rbpf_get_match(self._h, pose.ctypes.data_as(_dp), cov.ctypes.data_as(_dp),
                                          score.ctypes.data_as(_dp), valid.ctypes.data_as(_ip),
                                          best.ctypes.data_as(_ip)))
        return dict(pose=pose, cov=cov, score=score, valid=valid.astype(bool), best=best)

    def set_match(self, pose, cov, valid):
        pose = np.ascontiguousarray(np.broadcast_to(np.asarray(pose, dtype=np.float64), (self.N, 3)))
        cov = np.ascontiguousarray(np.broadcast_to(np.asarray(cov, dtype=np.float64), (self.N, 3, 3)))
        valid = np.ascontiguousarray(np.broadcast_to(np.asarray(valid, dtype=np.int32), (self.N,)))
        self._ck(self._lib.rbpf_set_match(self._h, pose.ctypes.data_as(_dp), cov.ctypes.data_as(_dp),
                                          valid.ctypes.data_as(_ip)))

    def match_slice(self, particle):
        out = np.zeros((29, 29), dtype=np.int32)
        self._ck(self._lib.rbpf_get_match_slice(self._h, int(particle), out.ctypes.data_as(_ip)))
        return out

    def list_tiles(self, particle):
        out = np.zeros((64, 2), dtype=np.int32)
        n = C.c_int32(0)
        self._ck(self._lib.rbpf_list_tiles(self._h, int(particle), out.ctypes.data_as(_ip), 64, C.byref(n)))
        return [tuple(int(v) for v in out[i]) for i in range(n.value)]

    def export_tile(self, particle, cx, cy):
        """800x800 float64 [ix][iy] like HybridMapEntry.map()._map, or None if absent."""
        out = np.empty((DIM, DIM), dtype=np.float64)
        ex = C.c_int32(0)
        self._ck(self._lib.rbpf_export_tile(self._h, int(particle), int(cx), int(cy), out.ctypes.data_as(_dp),
                                            C.byref(ex)))
        return out if ex.value else None

    def stats(self):
        s = _lib.RbpfStats()
        self._ck(self._lib.rbpf_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def weights_device_ptr(self):
        p = C.c_uint64(0)
        self._ck(self._lib.rbpf_weights_device_ptr(self._h, C.byref(p)))
        return p.value
