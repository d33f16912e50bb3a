"""Particle sharding across the GPUs of one box.

One process per GPU (torchrun), each owning a contiguous slice of the particles,
its own tile pool and page tables.  Stages 1-4 touch only a rank's own
particles; the one exchange step is resampling:

  1. all-gather of the weights over NCCL (8 B x N),
  2. every rank runs the identical global systematic resample
     (`rbpf_resample_global`, bit-exact ancestors on all ranks),
  3. survivors whose ancestor lives on another rank migrate.  Transport "peer"
     (default): every rank maps its peers' particle state once with CUDA IPC and
     then PULLS what it needs -- page tables, poses and de-duplicated sub-tiles --
     through NVLink straight into its own pool (`rbpf_migrate_pull`), followed by a
     one-element all-reduce as the "everybody has read" barrier -- on a side stream:
     only the reference-count pass that frees tiles waits for it, at the next scan's
     map integration (`rbpf_resample_apply_local_deferred`); no packing, no
     size negotiation, no host synchronisation in the data path.  Transport "nccl"
     (fallback, `RBPF_DIST_TRANSPORT=nccl`): the sender packs
     (`rbpf_migrate_count/pack`), NCCL point-to-point moves the buffers, the
     receiver adopts them (`rbpf_migrate_unpack`).

`plan_migration` is pure numpy (same result on every rank, no negotiation) and is
what the gloo CPU tests exercise; the byte movement is torch.distributed.
"""
import ctypes as C
import os

import numpy as np

from .particles import ParticleSet

_ip = C.POINTER(C.c_int32)


def owner_of(global_index, n_local):
    return global_index // n_local


def plan_migration(ancestors, n_local, rank, world, want_sources=False):
    """From the global ancestor vector derive, for this rank:
      send[r]  = sorted unique LOCAL slots whose particle rank r needs,
      recv[r]  = (dst_slots, rec_index, n_records): local destination slots fed by
                 rank r and, for each, the index into r's send list (== r's send[rank]).
    Every rank computes both sides from the same vector, so sizes agree.  The
    ancestor vector is non-decreasing (systematic resampling), so the slots fed by
    this rank's particles form one contiguous interval: O(n_local) work.
    want_sources: also return src[r] = the sorted unique slots of rank r this rank
    needs (what a pulling receiver reads; equals rank r's send[rank])."""
    anc = np.asarray(ancestors)
    lo_id, hi_id = rank * n_local, (rank + 1) * n_local
    empty = np.zeros(0, dtype=np.int32)
    send = {r: empty for r in range(world) if r != rank}
    recv = {r: (empty, empty, 0) for r in range(world) if r != rank}
    src_of = {r: empty for r in range(world) if r != rank}

    def uniq_sorted(a):
        """(unique values, index of every element in them) of a non-decreasing array, without sorting."""
        first = np.empty(len(a), dtype=bool)
        first[0] = True
        np.not_equal(a[1:], a[:-1], out=first[1:])
        return a[first], np.cumsum(first) - 1

    # what leaves: global slots j in [j0, j1) have an ancestor among my particles; the part of that
    # interval that lies in rank r's slice goes to r
    j0, j1 = int(np.searchsorted(anc, lo_id, "left")), int(np.searchsorted(anc, hi_id, "left"))
    for r in range(j0 // n_local, (j1 - 1) // n_local + 1 if j1 > j0 else 0):
        if r == rank:
            continue
        a, b = max(j0, r * n_local), min(j1, (r + 1) * n_local)
        if b > a:
            send[r] = (uniq_sorted(anc[a:b])[0] - lo_id).astype(np.int32)
    # what arrives: my slots whose ancestor lives elsewhere (mine is non-decreasing: one interval per source rank)
    mine = anc[lo_id:hi_id]
    cuts = np.searchsorted(mine, np.arange(world + 1) * n_local, "left")
    for r in range(world):
        a, b = int(cuts[r]), int(cuts[r + 1])
        if r == rank or b <= a:
            continue
        uniq, idx = uniq_sorted(mine[a:b])
        recv[r] = (np.arange(a, b, dtype=np.int32), idx.astype(np.int32), len(uniq))
        src_of[r] = (uniq - r * n_local).astype(np.int32)
    if want_sources:
        return send, recv, src_of
    return send, recv


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view library memory."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class MigratingSet(ParticleSet):
    """ParticleSet slice of one rank + the two halves of the resample exchange,
    independent of how the bytes travel (NCCL in ShardedParticleSet, plain device
    copies in the single-GPU emulation used by the tests)."""

    def __init__(self, n_local, n_beams, rank, world, device=0, **kw):
        import torch

        self._torch = torch
        super().__init__(n_local, n_beams, device=device, rank=rank, world=world, **kw)
        self.n_global = n_local * world
        self._dev = torch.device("cuda", device)
        self._w_local = torch.as_tensor(_DevArray(self.weights_device_ptr(), n_local, "<f8"), device=self._dev)
        self._anc = np.empty(self.n_global, dtype=np.int32)
        self.migrated_particles = 0
        self.migrated_bytes = 0
        self._arena = {}                  # (kind, peer) -> persistent uint8 buffer, grown geometrically

    def _buffer(self, kind, peer, nbytes):
        """Persistent staging buffer: sizes change every step and a fresh torch.empty
        of tens of MB can fall through to cudaMalloc (milliseconds)."""
        buf = self._arena.get((kind, peer))
        if buf is None or buf.numel() < nbytes:
            buf = self._torch.empty(max(int(nbytes * 1.5), 1 << 20), dtype=self._torch.uint8, device=self._dev)
            self._arena[(kind, peer)] = buf
        return buf[:nbytes]

    def local_weights_tensor(self):
        return self._w_local

    def pack_outgoing(self, weights_all, u01=None):
        """Global resample on the gathered weights, then pack what leaves this rank.
        Returns (did_resample, {peer: (buffer, n_particles, n_subtiles)}, recv_plan)."""
        lib, torch = self._lib, self._torch
        did = C.c_int32(0)
        up = None
        if u01 is not None:
            u = C.c_double(float(u01))
            up = C.cast(C.byref(u), C.POINTER(C.c_double))
        self._ck(lib.rbpf_resample_global(self._h, weights_all.data_ptr(), self.n_global, up, C.byref(did),
                                          self._anc.ctypes.data_as(_ip)))
        send, recv = plan_migration(self._anc, self.N, self.rank, self.world)
        out = {}
        for r, slots in send.items():
            if len(slots) == 0:                               # nothing for this peer: no device round trip
                out[r] = (None, 0, 0)
                continue
            nt, nbytes = C.c_int32(0), C.c_int64(0)
            self._ck(lib.rbpf_migrate_count(self._h, slots.ctypes.data_as(_ip), len(slots), C.byref(nt), C.byref(nbytes)))
            buf = self._buffer("send", r, int(nbytes.value))
            self._ck(lib.rbpf_migrate_pack(self._h, buf.data_ptr()))
            out[r] = (buf, len(slots), int(nt.value))
            self.migrated_particles += len(slots)
            self.migrated_bytes += buf.numel()
        return bool(did.value), out, recv

    # -- pull transport
    def peer_view(self):
        """Bytes describing this rank's buffers (CUDA IPC handles + raw pointers)."""
        from . import _lib as L

        v = L.RbpfPeerView()
        self._ck(self._lib.rbpf_peer_export(self._h, C.byref(v)))
        return bytes(v)

    def attach_peer(self, peer_rank, view_bytes):
        from . import _lib as L

        v = L.RbpfPeerView.from_buffer_copy(view_bytes)
        self._ck(self._lib.rbpf_peer_attach(self._h, int(peer_rank), C.byref(v)))

    def plan_and_pull(self, weights_all, u01=None, check=True):
        """Global resample on the gathered weights, then pull every remote ancestor this
        rank needs through the peer mappings.  The plan is the ancestor vector on the
        device: with check=False nothing is copied to the host and nothing waits (errors
        surface at the next synchronising call); check=True also fetches the ancestors and
        the reference's resample assertion.  The caller must make sure every rank has
        finished pulling before any rank calls finish_resample()."""
        lib = self._lib
        up = None
        if u01 is not None:
            u = C.c_double(float(u01))
            up = C.cast(C.byref(u), C.POINTER(C.c_double))
        did = None
        if check:
            d = C.c_int32(0)
            self._ck(lib.rbpf_resample_global(self._h, weights_all.data_ptr(), self.n_global, up, C.byref(d),
                                              self._anc.ctypes.data_as(_ip)))
            did = bool(d.value)
            mine = self._anc[self.rank * self.N:(self.rank + 1) * self.N] // self.N
            self.migrated_particles += int(np.count_nonzero(mine != self.rank))
        else:
            self._ck(lib.rbpf_resample_global(self._h, weights_all.data_ptr(), self.n_global, up, None, None))
        prof = getattr(self, "_prof", None)
        if prof is not None:                                  # RBPF_DIST_PROFILE: the plan and the pull apart (adds a sync)
            import time

            self._torch.cuda.synchronize()
            t0 = time.perf_counter()
        # sharded jobs (ShardedParticleSet): the sub-tile payloads travel on the side stream, in front of the barrier;
        # the emulations of the tests pull in stream order
        side = getattr(self, "_side", None) if getattr(self, "_overlap_pull", False) else None
        if side is not None:
            self._ck(lib.rbpf_migrate_pull_async(self._h, int(side.cuda_stream)))
        else:
            self._ck(lib.rbpf_migrate_pull(self._h))
        if prof is not None:
            self._torch.cuda.synchronize()
            prof["pull"] = prof.get("pull", 0.0) + time.perf_counter() - t0
        return did

    def finish_resample(self):
        self._ck(self._lib.rbpf_resample_apply_local(self._h))
        self._ck(self._lib.rbpf_resample_commit(self._h))

    def adopt_incoming(self, incoming, recv_plan):
        """Local gather / refcounts, then adopt {peer: (buffer, n_particles, n_subtiles)}."""
        lib = self._lib
        self._ck(lib.rbpf_resample_apply_local(self._h))
        for r, (buf, n_in, t_in) in incoming.items():
            dst_slots, rec_idx, n_plan = recv_plan[r]
            assert n_in == n_plan, "migration plan mismatch between ranks"
            if n_in:
                self._ck(lib.rbpf_migrate_unpack(self._h, buf.data_ptr(), n_in, t_in, dst_slots.ctypes.data_as(_ip),
                                                 rec_idx.ctypes.data_as(_ip), len(dst_slots)))
        self._ck(lib.rbpf_resample_commit(self._h))

    def recv_bytes(self, n_particles, n_subtiles):
        return int(self._lib.rbpf_migrate_bytes(self._h, n_particles, n_subtiles))


class ShardedParticleSet(MigratingSet):
    """One rank of a torch.distributed (NCCL) job."""

    def __init__(self, n_local, n_beams, group=None, device=0, **kw):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        if kw.get("stream", 0):
            # the collectives below are ordered against the handle's work through the legacy default
            # stream; a private stream would let peers read tiles that finish_resample() is freeing
            raise ValueError("ShardedParticleSet runs on the default stream (stream=0)")
        super().__init__(n_local, n_beams, dist.get_rank(group), dist.get_world_size(group), device=device, **kw)
        self._w_all = self._torch.empty(self.n_global, dtype=self._torch.float64, device=self._dev)
        self._tev = None
        self.transport = os.environ.get("RBPF_DIST_TRANSPORT", "peer")
        if self.transport not in ("peer", "nccl"):
            raise ValueError("RBPF_DIST_TRANSPORT must be 'peer' or 'nccl'")
        if self.transport == "peer" and self.world > 16:       # RB_MAX_WORLD peer mappings per kernel launch
            self.transport = "nccl"
        if self.transport == "peer":
            self._attach_all()
        self._prof = {} if os.environ.get("RBPF_DIST_PROFILE") else None   # host-side phase timers (adds syncs)

    def _attach_all(self):
        """Exchange the buffer descriptions and map every peer (once per job).  All ranks must
        agree on the transport, so a failure anywhere switches every rank back to NCCL."""
        torch, dist = self._torch, self._dist
        mine = torch.frombuffer(bytearray(self.peer_view()), dtype=torch.uint8).to(self._dev)
        views = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self._dev)
        dist.all_gather_into_tensor(views, mine, group=self.group)
        views = views.cpu().numpy().reshape(self.world, -1)
        ok = 1
        why = ""
        for r in range(self.world):
            if r == self.rank:
                continue
            try:
                self.attach_peer(r, views[r].tobytes())
            except Exception as e:               # no IPC / no peer access on this box
                ok, why = 0, str(e)
                break
        flag = torch.tensor([ok], dtype=torch.int32, device=self._dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            import sys

            if self.rank == 0 or why:
                print("[rbpf] rank %d: peer transport unavailable (%s); every rank uses NCCL send/recv" % (self.rank, why or "a peer failed"),
                      file=sys.stderr)
            self.transport = "nccl"
        self._barrier_flag = torch.zeros(1, dtype=torch.int32, device=self._dev)
        self._defer_barrier = os.environ.get("RBPF_DIST_BARRIER", "deferred") != "inline"
        # RBPF_DIST_OVERLAP=1: payload copies of the pull beside the rest of the resample and the next scan's matching (needs
        # the deferred barrier: the barrier has to follow the copies on the side stream).  Off by default: measured at 4 GPUs
        # it gains nothing (the copy takes CTA slots from the matcher, the migrated particles are matched in a second launch)
        self._overlap_pull = self._defer_barrier and os.environ.get("RBPF_DIST_OVERLAP", "0") == "1"
        self._side = torch.cuda.Stream(device=self._dev, priority=-1)
        self._gate = torch.cuda.Event()

    def _resample_peer(self, u01, want_ancestors):
        import time

        torch, dist = self._torch, self._dist
        prof = self._prof
        t0 = time.perf_counter()
        dist.all_gather_into_tensor(self._w_all, self._w_local, group=self.group)
        if prof is not None:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        did = self.plan_and_pull(self._w_all, u01, check=want_ancestors)
        if prof is not None:
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        if self._defer_barrier:
            # "every rank has read what it needs" off the critical path: the barrier runs on a side stream behind the
            # pull, the gather of the local ancestors runs now, and the reference counts -- frees and in-place writes
            # of tiles a peer may still be reading -- wait for the barrier's event at the next call that needs the
            # pool (the next scan's integrate): motion, matching and weighting of that scan only read tiles
            side = self._side
            side.wait_stream(torch.cuda.current_stream(self._dev))
            with torch.cuda.stream(side):
                dist.all_reduce(self._barrier_flag, group=self.group)
                self._gate.record(side)
            if prof is not None:
                torch.cuda.synchronize()
                t3 = time.perf_counter()
            self._ck(self._lib.rbpf_resample_apply_local_deferred(self._h, int(self._gate.cuda_event)))
            self._ck(self._lib.rbpf_resample_commit(self._h))
        else:
            dist.all_reduce(self._barrier_flag, group=self.group)        # every rank has read what it needs
            if prof is not None:
                torch.cuda.synchronize()
                t3 = time.perf_counter()
            self.finish_resample()
        if prof is not None:
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            for k, v in (("allgather", t1 - t0), ("plan+pull", t2 - t1), ("barrier", t3 - t2), ("apply", t4 - t3)):
                prof[k] = prof.get(k, 0.0) + v
            prof["n"] = prof.get("n", 0) + 1
        return did, (self._anc.copy() if want_ancestors else None)

    def resample(self, u01=None, want_ancestors=True):
        # NCCL work is enqueued on torch's CURRENT stream: pin it to the default stream, which is the
        # handle's, whatever torch.cuda.stream(...) context the caller is in
        with self._torch.cuda.stream(self._torch.cuda.default_stream(self._dev)):
            if self.transport == "peer":
                return self._resample_peer(u01, want_ancestors)
            return self._resample_nccl(u01, want_ancestors)

    def _resample_nccl(self, u01, want_ancestors):
        import time

        torch, dist = self._torch, self._dist
        prof = self._prof
        t0 = time.perf_counter()
        dist.all_gather_into_tensor(self._w_all, self._w_local, group=self.group)
        if prof is not None:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        did, out, recv = self.pack_outgoing(self._w_all, u01)
        if prof is not None:
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        meta = torch.zeros((self.world, 2), dtype=torch.int64)
        for r, (_, n, nt) in out.items():
            meta[r, 0], meta[r, 1] = n, nt
        meta_dev = meta.to(self._dev)
        meta_in = torch.empty_like(meta_dev)
        dist.all_to_all_single(meta_in, meta_dev, group=self.group)       # sub-tile counts of what arrives
        meta_in = meta_in.cpu()
        incoming, ops = {}, []
        for r in range(self.world):
            if r == self.rank:
                continue
            n_in, t_in = int(meta_in[r, 0]), int(meta_in[r, 1])
            if n_in:
                buf = self._buffer("recv", r, self.recv_bytes(n_in, t_in))
                incoming[r] = (buf, n_in, t_in)
                ops.append(dist.P2POp(dist.irecv, buf, r, group=self.group))
            else:
                incoming[r] = (None, 0, 0)
            if out[r][1]:
                ops.append(dist.P2POp(dist.isend, out[r][0], r, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if prof is not None:
            torch.cuda.synchronize()
            t3 = time.perf_counter()
        self.adopt_incoming(incoming, recv)
        if prof is not None:
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            for k, v in (("allgather", t1 - t0), ("plan+pack", t2 - t1), ("exchange", t3 - t2), ("adopt", t4 - t3)):
                prof[k] = prof.get(k, 0.0) + v
            prof["n"] = prof.get("n", 0) + 1
        return did, (self._anc.copy() if want_ancestors else None)

    def step(self, ranges, angles):
        """One lidar event on this rank's slice + the global resample."""
        ev = self._tev
        marks = []

        def mark():
            if ev is not None:
                e = self._torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)

        mark(); self.set_scan(ranges, angles)
        mark(); self.scan_match()
        mark(); self.weight(None)
        mark(); self.integrate(fallback_weights=True)
        mark(); self.resample(None, want_ancestors=False)
        mark()
        if ev is not None:
            ev.append(marks)

    def timing_enable(self, max_steps):
        self._tev = [] if max_steps > 0 else None
        if self._prof is not None and max_steps > 0:
            self._prof.clear()
            self.migrated_particles = self.migrated_bytes = 0

    def timing_read(self):
        out = {k: 0.0 for k in self.STAGES}
        steps = self._tev or []
        self._torch.cuda.synchronize()
        for m in steps:
            out["set_scan"] += m[0].elapsed_time(m[1])
            out["match"] += m[1].elapsed_time(m[2])
            out["weight"] += m[2].elapsed_time(m[3])
            out["raycast_cast"] += m[3].elapsed_time(m[4])       # prepare + cast + fallback weights
            out["resample_plan"] += m[4].elapsed_time(m[5])      # all-gather + plan + migration + apply
        n = len(steps)
        if self._tev is not None:
            self._tev = []
        return out, n


def dist_parity_check(particles_per_rank=1024, n_beams=360, frames=8, device=0, seed=5, tmp_dir=None):
    """configs[3] in small: a Freiburg-shaped CARMEN log (360 beams; the real fr.log is missing from the
    reference tree, so thesis_b200.synth writes a stand-in), `particles_per_rank` particles on every rank,
    driven through the drop-in API -- `[Robot(eng) ...]`, the headless main.py loop, `resample` -- once
    sharded over all ranks of the job (NCCL all-gather of the weights, NVLink migration) and once as a
    single set on this rank's GPU.  Every rank compares its slice: poses, covariances, weights and the
    maps of a few of its particles must be identical bit for bit.  Collective: call on every rank.
    Returns {ranks, particles, particles_per_rank, frames, migrated, identical}."""
    import tempfile

    import torch
    import torch.distributed as dist

    from . import harness, loaders, particles as P, sensors, synth

    rank, world = dist.get_rank(), dist.get_world_size()
    n = particles_per_rank * world
    dev = torch.device("cuda", device)
    own_tmp = None
    if tmp_dir is None:
        own_tmp = tempfile.TemporaryDirectory(prefix="rbpf_parity_r%d_" % rank)
        tmp_dir = own_tmp.name
    w = synth.Workload(frames + 1, n_beams=n_beams)                 # every rank writes its own copy of the same log
    synth.write_carmen_log(os.path.join(tmp_dir, "fr.txt"), w)
    synth.write_carmen_log(os.path.join(tmp_dir, "fr.log"), w)

    def run(sharded):
        ld = sensors.Lidar(loaders.FreidLidarData(tmp_dir))
        im = sensors.IMU(loaders.FreidIMUData(tmp_dir))
        sh = P.new_filter(rng="device", keep_history=False, sharded=sharded, device=device, seed=seed, world_tiles=(5, 5),
                          pool_subtiles=40 * (particles_per_rank if sharded else n) + 4096)
        parts = [P.Robot("eng") for _ in range(n)]                   # main.py:87 on every rank
        migrated = 0

        def resample(ps):
            nonlocal migrated
            out = P.resample(ps)
            anc = ps[0]._shared.last_ancestors
            if sharded:
                lo = rank * particles_per_rank
                mine = anc[lo:lo + particles_per_rank] // particles_per_rank
                migrated += int(np.count_nonzero(mine != rank))
            return out

        harness.run_log(parts, ld, im, resample, seed_fn=P.seed_map, max_frames=frames)
        sh.ps.synchronize()
        return sh, migrated

    sh_a, migrated = run(True)
    sh_b, _ = run(False)
    a, b = sh_a.ps, sh_b.ps
    lo, hi = rank * particles_per_rank, (rank + 1) * particles_per_rank
    same = (np.array_equal(a.poses, b.poses[lo:hi]) and np.array_equal(a.covs, b.covs[lo:hi]) and
            np.array_equal(a.weights, b.weights[lo:hi]))
    for j in (0, particles_per_rank // 2, particles_per_rank - 1):
        same = same and sorted(a.list_tiles(j)) == sorted(b.list_tiles(lo + j))
        for c in a.list_tiles(j):
            same = same and bool(np.array_equal(a.export_tile(j, *c), b.export_tile(lo + j, *c)))
    t = torch.tensor([1 if same else 0, -migrated], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = bool(int(t[0]))
    t2 = torch.tensor([migrated], dtype=torch.int64, device=dev)
    dist.all_reduce(t2, op=dist.ReduceOp.SUM)
    a.close()
    b.close()
    P._DEFAULT_SET = None
    if own_tmp is not None:
        own_tmp.cleanup()
    return {"ranks": world, "particles": n, "particles_per_rank": particles_per_rank, "beams": n_beams, "frames": frames,
            "log": "synthetic CARMEN log through FreidLidarData / FreidIMUData, drop-in Robot / resample API",
            "migrated": int(t2[0]), "identical": ok, "transport": a.transport}
