"""GPU: the sharded resample (global plan + sub-tile migration) emulated with two
rank slices on ONE device -- buffers are handed over directly instead of through
NCCL -- must reproduce a single set holding all particles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from thesis_b200 import dist, particles

    return torch, dist, particles


@pytest.mark.parametrize("transport,world", [("nccl", 2), ("peer", 2), ("peer", 3), ("peer-async", 3)])
def test_two_rank_emulation_equals_single_set(mods, golden, transport, world):
    """Ranks emulated on one GPU.  "nccl": pack -> (copy) -> unpack; "peer": every rank
    pulls what it needs straight out of the other ranks' buffers (same-process mapping
    instead of CUDA IPC), then all ranks apply and commit.  "peer-async": the payload copies of
    the pull run on a side stream per rank (rbpf_migrate_pull_async), the barrier is an event behind
    all of them, the reference-count pass is deferred behind that event
    (rbpf_resample_apply_local_deferred) and the next scan's matcher takes the local particles first
    and the migrated ones behind the copies.  Either way the union of the ranks must equal one
    ParticleSet bit for bit, every step."""
    torch, D, P = mods
    nl, B, K = 6, 180, 30
    N = world * nl
    ang = golden["intel_angles"]
    rng = np.random.default_rng(21)
    one = P.ParticleSet(N, B, pool_subtiles=3000)
    ranks = [D.MigratingSet(nl, B, r, world, pool_subtiles=2000) for r in range(world)]
    sides = []
    if transport == "peer-async":
        sides = [torch.cuda.Stream() for _ in ranks]
        bar, gate = torch.cuda.Stream(), torch.cuda.Event()
        for ps, side in zip(ranks, sides):
            ps._side, ps._overlap_pull = side, True
    if transport.startswith("peer"):
        views = [ps.peer_view() for ps in ranks]
        for k, ps in enumerate(ranks):
            for r_ in range(world):
                if r_ != k:
                    ps.attach_peer(r_, views[r_])
    par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
    r0 = golden["intel_ranges"][0]
    for ps in [one] + ranks:
        for _ in range(2):
            ps.set_scan(r0, ang)
            ps.integrate()
    for step in range(1, 6):
        u = (rng.normal(0.05, 0.02), rng.normal(0, 0.02), rng.normal(-0.3, 0.05))
        z = rng.standard_normal((N, K, 3))
        u01 = float(rng.random())
        r = golden["intel_ranges"][step]
        one.motion(1, u, 1.0, par)
        one.set_scan(r, ang); one.scan_match(); one.weight(z); one.integrate(fallback_weights=True)
        did1, anc1 = one.resample(u01)
        for k, ps in enumerate(ranks):
            ps.motion(1, u, 1.0, par)
            ps.set_scan(r, ang); ps.scan_match(); ps.weight(z[k * nl:(k + 1) * nl]); ps.integrate(fallback_weights=True)
        w_all = torch.cat([ps.local_weights_tensor().clone() for ps in ranks])       # the all-gather
        if transport.startswith("peer"):
            for ps in ranks:
                assert ps.plan_and_pull(w_all, u01) == did1
                assert np.array_equal(ps._anc, anc1), "step %d: ancestors differ from the single set" % step
            if sides:
                for side in sides:                                                   # the barrier: an event behind every rank's copies
                    bar.wait_stream(side)
                gate.record(bar)
                for ps in ranks:
                    ps._ck(ps._lib.rbpf_resample_apply_local_deferred(ps._h, int(gate.cuda_event)))
                    ps._ck(ps._lib.rbpf_resample_commit(ps._h))
            else:
                torch.cuda.synchronize()                                             # the barrier
                for ps in ranks:
                    ps.finish_resample()
        else:
            packed = [ps.pack_outgoing(w_all, u01) for ps in ranks]
            for k, ps in enumerate(ranks):
                assert packed[k][0] == did1
                assert np.array_equal(ps._anc, anc1), "step %d: ancestors differ from the single set" % step
            for k, ps in enumerate(ranks):
                incoming = {r_: packed[r_][1][k] for r_ in range(world) if r_ != k}  # the all-to-all
                ps.adopt_incoming(incoming, packed[k][2])
        poses = np.concatenate([ps.poses for ps in ranks])
        assert np.array_equal(poses, one.poses), "step %d" % step
        assert np.array_equal(np.concatenate([ps.weights for ps in ranks]), one.weights)
        assert np.array_equal(np.concatenate([ps.covs for ps in ranks]), one.covs)
    assert sum(ps.migrated_particles for ps in ranks) > 0, "test never exercised a migration"
    for j in range(N):
        ps, lj = ranks[j // nl], j % nl
        assert sorted(ps.list_tiles(lj)) == sorted(one.list_tiles(j))
        for (cx, cy) in one.list_tiles(j):
            assert np.array_equal(ps.export_tile(lj, cx, cy), one.export_tile(j, cx, cy)), "particle %d tile %s" % (j, (cx, cy))
    # pool accounting: nothing leaked on either rank
    for ps in ranks:
        st = ps.stats()
        assert st["pool_in_use"] <= st["total_refs"]


def test_sharded_drop_in_api_world_of_one(mods, tmp_path):
    """`new_filter(sharded=True)` + `[Robot(eng) ...]` + the headless loop + `resample` under a real
    (one-rank) NCCL process group: the sharded path of the drop-in API -- all-gathered poses for
    particle 0's gate, global resample -- gives the same result as the plain single set.
    (profiles/ holds the multi-rank runs of thesis_b200.dist.dist_parity_check.)"""
    torch, D, P = mods
    import torch.distributed as dist

    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        r = D.dist_parity_check(particles_per_rank=96, n_beams=360, frames=6, device=0, tmp_dir=str(tmp_path))
        assert r["identical"] and r["ranks"] == 1 and r["particles"] == 96 and r["migrated"] == 0
        with pytest.raises(P.RbpfError):
            P.new_filter(rng="numpy", sharded=True)                # host draws cannot be sharded
    finally:
        dist.destroy_process_group()


def test_snapshot_restore_rewinds_the_filter(mods, golden):
    """rbpf_snapshot / rbpf_restore: the same scans from the same snapshot give the same particles
    (device draws are keyed by the step counter, which is restored too)."""
    torch, D, P = mods
    ps = P.ParticleSet(64, 180, pool_subtiles=6000, seed=9)
    r, a = golden["intel_ranges"], golden["intel_angles"]
    ps.set_scan(r[0], a); ps.integrate(); ps.integrate()
    par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
    for s in range(1, 4):
        ps.motion(1, (0.05, 0.0, -0.3), 1.0, par); ps.step(r[s], a)
    ps.snapshot()

    def run():
        for s in range(4, 9):
            ps.motion(1, (0.05, 0.0, -0.3), 1.0, par); ps.step(r[s], a)
        ps.synchronize()
        return ps.poses.copy(), ps.weights.copy(), ps.export_tile(5, 0, 0).copy(), ps.stats()

    p1, w1, t1, s1 = run()
    ps.restore()
    p2, w2, t2, s2 = run()
    assert np.array_equal(p1, p2) and np.array_equal(w1, w2) and np.array_equal(t1, t2)
    assert s1["pool_in_use"] == s2["pool_in_use"] and s2["refcount_sum"] == s2["total_refs"]
