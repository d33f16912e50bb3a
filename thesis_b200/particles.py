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
                 device=0, rank=0, world=1, stream=0, seed=0, ndt_refine=False):
        """ndt_refine: run the NDT stage of the reference matcher (matchScanCustom.m:32-50) after the grid search."""
        self._lib = _lib.load()
        self.N, self.B, self.K = int(n_particles), int(n_beams), int(n_samples)
        self.rank, self.world = rank, world
        if pool_subtiles is None:
            pool_subtiles = max(4096, 48 * self.N)
        cfg = _lib.RbpfConfig(self.N, self.B, self.K, world_tiles[0], world_tiles[1], int(pool_subtiles),
                              device, rank, world, 1 if ndt_refine else 0, int(stream), int(seed))
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

    def snapshot(self):
        """Copy the whole mutable state into device-side shadow buffers (doubles the memory on first use)."""
        self._ck(self._lib.rbpf_snapshot(self._h))

    def restore(self):
        """Rewind to the last snapshot()."""
        self._ck(self._lib.rbpf_restore(self._h))

    def clear_errors(self):
        """Reset the sticky device-side error flags (pool exhausted, resample assertion, internal bound)."""
        self._ck(self._lib.rbpf_clear_errors(self._h))

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

    def scan_match(self, last_scan_xy=None):
        """Scan-to-map, or scan-to-previous-scan when last_scan_xy ([n, 2] global endpoints) is given."""
        if last_scan_xy is None:
            self._ck(self._lib.rbpf_scan_match(self._h))
        else:
            xy, p = _d(last_scan_xy)
            self._ck(self._lib.rbpf_scan_match_adj(self._h, p, xy.shape[0]))

    def weight(self, z=None):
        if z is None:
            self._ck(self._lib.rbpf_weight(self._h, None))
        else:
            z, zp = _d(z)
            if z.size != self.N * self.K * 3:
                raise ValueError("z must hold N*K*3 standard normals")
            self._ck(self._lib.rbpf_weight(self._h, zp))

    def weight_guesses(self, guesses):
        """Weight stage on caller-drawn proposal samples [N, K, 3] (robot.py:81 drawn with NumPy itself)."""
        g, gp = _d(guesses)
        if g.size != self.N * self.K * 3:
            raise ValueError("guesses must hold N*K*3 values")
        self._ck(self._lib.rbpf_weight_guesses(self._h, gp))

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
        refine = np.empty((self.N, 2), dtype=np.int32)
        self._ck(self._lib.rbpf_get_match_refine(self._h, refine.ctypes.data_as(_ip)))
        return dict(pose=pose, cov=cov, score=score, valid=valid.astype(bool), best=best,
                    ndt_evals=refine[:, 0].copy(), ndt_accepted=refine[:, 1].astype(bool))

    def resample_cumsum(self):
        """Running sum of the adjusted weights of the last triggered resample (main.py:57,62)."""
        out = np.empty(self.N * self.world)
        self._ck(self._lib.rbpf_get_resample_cumsum(self._h, out.ctypes.data_as(_dp)))
        return out

    def set_refine(self, on):
        """Switch the NDT stage (matchScanCustom.m:32-50) on or off for the following matches."""
        self._ck(self._lib.rbpf_set_refine(self._h, 1 if on else 0))

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

    def occupied_points(self, particle):
        """[n, 2] cell coordinates of cells with log-odds > 1.0 (hybridmap.py:303-313), compacted on the device."""
        n = C.c_int64(0)
        self._ck(self._lib.rbpf_occupied_points(self._h, int(particle), None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 2), dtype=np.float64)
        self._ck(self._lib.rbpf_occupied_points(self._h, int(particle), out.ctypes.data_as(_dp), n.value, C.byref(n)))
        return out[: n.value]

    def save(self, path):
        """Checkpoint the whole set (the reference shelves particle 0 only, main.py:183-210)."""
        self._ck(self._lib.rbpf_checkpoint_write(self._h, str(path).encode()))

    def load(self, path):
        self._ck(self._lib.rbpf_checkpoint_read(self._h, str(path).encode()))

    def stats(self):
        s = _lib.RbpfStats()
        self._ck(self._lib.rbpf_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    MATCH_PHASES = ("frame_points", "gather", "dilations", "ref_mask", "seeds_bounds", "rank", "members",
                    "covariance", "ndt")

    def match_phase_clocks(self):
        """SM clocks of the matcher kernel by phase (rbpf_match_phase_clocks): CTA-level split, per-warp
        busy clocks of the three search phases and the points they visited."""
        out = (C.c_uint64 * 16)()
        self._ck(self._lib.rbpf_match_phase_clocks(self._h, out))
        v = [int(x) for x in out]
        d = {"cta_" + n: v[i] for i, n in enumerate(self.MATCH_PHASES)}
        d.update(warp_seeds=v[10], warp_bounds=v[11], warp_members=v[12],
                 visits_seeds=v[13], visits_bounds=v[14], visits_members=v[15])
        return d

    def weights_device_ptr(self):
        p = C.c_uint64(0)
        self._ck(self._lib.rbpf_weights_device_ptr(self._h, C.byref(p)))
        return p.value


# ---------------------------------------------------------------------------
# The reference's particle API on top of the device-resident set.
# ---------------------------------------------------------------------------

class _SharedSet:
    """State shared by all Robot views of one filter."""

    def __init__(self, rng="numpy", keep_history=True, sharded=False, group=None, **set_kwargs):
        self.views = []
        self.ps = None
        self.rng = rng                    # "numpy": draws from np.random like the reference; "device": Philox
        self.keep_history = keep_history
        self.set_kwargs = set_kwargs
        # sharded: one process per GPU (torch.distributed, NCCL); every rank builds the SAME list of
        # Robot views -- main.py:87 unchanged -- and owns the slice [rank * n, (rank + 1) * n) of it
        self.sharded = bool(sharded)
        self.group = group
        self.rank, self.world, self.n_local = 0, 1, None
        if self.sharded:
            if rng != "device" or keep_history:
                raise RbpfError("a sharded filter draws on the device and keeps no per-view histories: "
                                "new_filter(rng='device', keep_history=False, sharded=True)")
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RbpfError("sharded=True needs an initialised torch.distributed process group (torchrun)")
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.motion_rounds = 0
        self.update_rounds = 0
        self._poses = None                # host cache of poses / weights, refreshed lazily
        self._weights = None
        self._covs = None
        self.last_ancestors = None

    def materialise(self, n_beams=384):
        if self.ps is None:
            kw = dict(self.set_kwargs)
            kw.setdefault("n_beams", n_beams)
            if self.sharded:
                from .dist import ShardedParticleSet

                if len(self.views) % self.world:
                    raise RbpfError("%d particles do not divide over %d ranks" % (len(self.views), self.world))
                self.n_local = len(self.views) // self.world
                nb = kw.pop("n_beams")
                self.ps = ShardedParticleSet(self.n_local, nb, group=self.group, **kw)
            else:
                self.n_local = len(self.views)
                self.ps = ParticleSet(len(self.views), **kw)
        return self.ps

    def invalidate(self):
        self._poses = self._weights = self._covs = None

    def _global(self, local):
        """Sharded: all-gather a per-particle array so that every rank sees every particle (main.py reads
        particle 0's pose on every frame, main.py:152-155,168).  A collective: every rank runs the same
        loop and therefore asks at the same point."""
        if not self.sharded:
            return local
        import torch
        import torch.distributed as dist

        ps = self.ps
        mine = torch.as_tensor(np.ascontiguousarray(local)).to(ps._dev)
        out = torch.empty((self.world,) + tuple(mine.shape), dtype=mine.dtype, device=ps._dev)
        with torch.cuda.stream(torch.cuda.default_stream(ps._dev)):
            dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.reshape((-1,) + tuple(mine.shape[1:])).cpu().numpy()

    def is_local(self, slot):
        return not self.sharded or self.n_local is None or self.rank * self.n_local <= slot < (self.rank + 1) * self.n_local

    def local_slot(self, slot):
        if not self.is_local(slot):
            raise RbpfError("particle %d lives on rank %d" % (slot, slot // self.n_local))
        return slot - self.rank * self.n_local if self.sharded else slot

    def poses(self):
        if self._poses is None:
            self._poses = self._global(self.materialise().poses)
        return self._poses

    def weights(self):
        if self._weights is None:
            self._weights = self._global(self.materialise().weights)
        return self._weights

    def covs(self):
        if self._covs is None:
            self._covs = self._global(self.materialise().covs)
        return self._covs

    def record_history(self):
        if not self.keep_history:
            return
        p, w = self.poses(), self.weights()
        for v in self.views:
            v._x.append(p[v._slot, 0])
            v._y.append(p[v._slot, 1])
            v._theta.append(p[v._slot, 2])
            if v._weight[-1] != w[v._slot]:
                v._weight.append(w[v._slot])


_DEFAULT_SET = None


def new_filter(rng="numpy", keep_history=True, sharded=False, group=None, **set_kwargs):
    """Start a new particle set; the Robot(...) constructions that follow join it.
    set_kwargs go to ParticleSet (world_tiles, pool_subtiles, device, seed, ...).

    sharded=True (under torchrun, one process per GPU, after torch.distributed.init_process_group("nccl")):
    the same `particles = [Robot(eng) for _ in range(NUM_PARTICLES)]` on every rank, but each rank's GPU
    holds NUM_PARTICLES / world of them (thesis_b200.dist.ShardedParticleSet); `resample(particles)` is
    the global systematic resample with NVLink migration.  Every rank sees every particle's pose and
    weight (all-gathered on demand), maps only of its own particles."""
    global _DEFAULT_SET
    _DEFAULT_SET = _SharedSet(rng=rng, keep_history=keep_history, sharded=sharded, group=group, **set_kwargs)
    return _DEFAULT_SET


class _MapView:
    """What main.py touches on `robot._map` (hybridmap.py:63-327): `_cell_size`,
    `get_occupied_points()`, `get_odds_at` / `get_pr_at`, `update`."""

    def __init__(self, robot):
        self._robot = robot
        self._cell_size = 0.05
        self._map_len_m = 40

    def _tiles(self):
        sh = self._robot._shared
        ps, j = sh.materialise(), sh.local_slot(self._robot._slot)
        return {c: ps.export_tile(j, c[0], c[1]) for c in ps.list_tiles(j)}

    def get_occupied_points(self):
        """Cell coordinates of cells with log-odds > 1.0 (hybridmap.py:303-313), thresholded and
        compacted on the device (the reference's O(tiles * 800^2) Python loop dominates its frame time)."""
        sh = self._robot._shared
        pts = sh.materialise().occupied_points(sh.local_slot(self._robot._slot))
        return list(pts[:, 0]), list(pts[:, 1])

    def get_odds_at(self, pos):
        for (cx, cy), t in self._tiles().items():
            if cx - 20.0 <= pos.x < cx + 20.0 and cy - 20.0 <= pos.y < cy + 20.0:
                return t[int((pos.x - cx) / 40 * 800 + 400)][int((pos.y - cy) / 40 * 800 + 400)]
        return None

    def get_pr_at(self, pos):
        v = self.get_odds_at(pos)
        if v is None:
            return None
        o = np.exp(v)
        return o / (1 + o)

    def __str__(self):
        sh = self._robot._shared
        return "Hybrid Map: %d maps" % len(sh.materialise().list_tiles(sh.local_slot(self._robot._slot)))


class Robot:
    """A particle (reference robot.py:19-157) as a view into the device set.

    `Robot(eng)` joins the current filter (see `new_filter`); `imu_update`,
    `map_update`, `weight`, `x`/`y`/`theta`, `get_latest_pose`, `copy` keep the
    reference's signatures.  The first view called in a round runs the batched
    CUDA kernels for every particle; the other calls of that round only catch up.
    """

    def __init__(self, matlab, _shared=None):
        if matlab is None and _shared is None:
            return                                        # Robot(None): the reference's empty shell (robot.py:20-21,142)
        global _DEFAULT_SET
        sh = _shared or _DEFAULT_SET or new_filter()
        if sh.ps is not None:
            raise RbpfError("the particle set is already on the device; call new_filter() before creating more Robots")
        self._shared = sh
        self._slot = len(sh.views)
        sh.views.append(self)
        self._map = _MapView(self)
        self._weight = [1.0]
        self._x, self._y, self._theta = [0.0], [0.0], [0.0]
        self._motion_seen = 0
        self._update_seen = 0

    # -- reference accessors (histories; with keep_history=False only the latest value is kept,
    #    read lazily from the device so that resample() costs no per-view Python work)
    def _latest(self, k):
        return [float(self._shared.poses()[self._slot, k])] if self._shared.ps is not None else [0.0]

    def x(self):
        return self._x if self._shared.keep_history else self._latest(0)

    def y(self):
        return self._y if self._shared.keep_history else self._latest(1)

    def theta(self):
        return self._theta if self._shared.keep_history else self._latest(2)

    def weight(self):
        if self._shared.keep_history or self._shared.ps is None:
            return self._weight
        return [float(self._shared.weights()[self._slot])]

    @property
    def _cov(self):
        return self._shared.covs()[self._slot]

    def get_latest_pose(self):
        p = self._shared.poses()[self._slot] if self._shared.ps is not None else (0.0, 0.0, 0.0)
        return Pose(float(p[0]), float(p[1]), float(p[2]))

    def __str__(self):
        return "Robot at position: " + str(self.get_latest_pose())

    # -- stage 1, robot.py:45-57
    def imu_update(self, reading):
        sh = self._shared
        if self._motion_seen == sh.motion_rounds:
            if reading.motion is None:
                raise RbpfError("this Reading carries no motion family; use the loaders in thesis_b200.loaders")
            family, par = reading.motion
            ps = sh.materialise()
            ps.motion(family, np.asarray(reading.get_data(), dtype=np.float64), reading.dt() / 1e4, par)
            sh.motion_rounds += 1
            sh.invalidate()
            sh.record_history()
        self._motion_seen += 1
        return self.get_latest_pose()

    # -- stages 2-4, robot.py:59-115
    def map_update(self, scan, last_scan=None, adj=False):
        sh = self._shared
        if self._update_seen == sh.update_rounds:
            if scan.ranges() is None:
                raise RbpfError("map_update needs a Scan built from ranges (Lidar[i])")
            ps = sh.materialise(len(scan))
            ps.set_scan(scan.ranges(), scan.angles())
            if adj:                  # robot.py:66-67 -> HybridMap.get_scan_adj hybridmap.py:147-191
                ps.scan_match(np.column_stack((last_scan.x(), last_scan.y())))
            else:                    # robot.py:68-69 -> HybridMap.get_scan_match hybridmap.py:210-261
                ps.scan_match()
            if sh.rng == "numpy":
                # robot.py:81 verbatim, with NumPy itself and in the reference's order (particle-major, no
                # draw for a failed match, robot.py:73-78): the global RNG stream is consumed exactly as
                # the reference consumes it and the samples are bit-identical to its samples
                m = ps.match_result()
                g = np.zeros((ps.N, ps.K, 3))
                for i in np.flatnonzero(m["valid"]):
                    g[i] = np.random.multivariate_normal(m["pose"][i], m["cov"][i], ps.K)
                ps.weight_guesses(g)
            else:
                ps.weight(None)
            ps.integrate(fallback_weights=True)
            sh.update_rounds += 1
            sh.invalidate()
            sh.record_history()
        self._update_seen += 1

    def copy(self):
        raise RbpfError("Robot.copy(): duplicates are made on the device by resample() (copy-on-write)")


def make_particles(n, rng="numpy", keep_history=True, **set_kwargs):
    """Replacement for `particles = [Robot(eng) for _ in range(NUM_PARTICLES)]` (main.py:87)."""
    sh = new_filter(rng=rng, keep_history=keep_history, **set_kwargs)
    return [Robot("gpu", _shared=sh) for _ in range(n)]


def seed_map(particles, scan, times=2):
    """The map seeding of main.py:89-90: integrate `scan` at every particle's pose."""
    sh = particles[0]._shared
    ps = sh.materialise(len(scan))
    ps.set_scan(scan.ranges(), scan.angles())
    for _ in range(times):
        ps.integrate()


def resample(particles):
    """main.resample (main.py:46-79): systematic resampling when max - min > 200.
    Returns the new particle list (views re-bound to the resampled slots)."""
    sh = particles[0]._shared
    ps = sh.materialise()
    if sh.rng == "numpy":
        w = sh.weights()
        # the reference draws its uniform only when the trigger fires (main.py:50,59)
        u01 = float(np.random.random()) if np.max(w) - np.min(w) > 200 else 0.5
        did, anc = ps.resample(u01)
    else:
        did, anc = ps.resample(None)
    sh.last_ancestors = anc
    sh.invalidate()
    if did:
        if sh.keep_history:
            old = [(list(v._x), list(v._y), list(v._theta), list(v._weight)) for v in sh.views]
            for v in sh.views:
                hx, hy, ht, hw = old[anc[v._slot]]
                v._x, v._y, v._theta, v._weight = list(hx), list(hy), list(ht), list(hw) + [1.0]
    return sh.views if not sh.keep_history else list(sh.views)
