"""Multi-process check of the sharded resample over real NCCL (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check_nccl.py

Every rank steps its slice with host-supplied draws; rank 0 additionally steps ONE
set holding all particles and compares after every scan: ancestors, poses,
covariances, weights and (at the end) the maps of every particle must be identical.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thesis_b200.dist import ShardedParticleSet  # noqa: E402
from thesis_b200.particles import ParticleSet  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    nl, B, K = 24, 180, 30
    N = nl * world
    ang = G["intel_angles"]
    rng = np.random.default_rng(5)
    sh = ShardedParticleSet(nl, B, device=lr, pool_subtiles=8000)
    one = ParticleSet(N, B, device=lr, pool_subtiles=8000 * world) if rank == 0 else None
    for ps in (sh, one):
        if ps is None:
            continue
        for _ in range(2):
            ps.set_scan(G["intel_ranges"][0], ang)
            ps.integrate()
    par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
    ok = True
    migrated = 0
    for step in range(1, 13):
        u = (rng.normal(0.05, 0.02), rng.normal(0, 0.02), rng.normal(-0.3, 0.05))
        z = rng.standard_normal((N, K, 3))
        u01 = float(rng.random())
        r = G["intel_ranges"][step]
        sh.motion(1, u, 1.0, par)
        sh.set_scan(r, ang); sh.scan_match(); sh.weight(z[rank * nl:(rank + 1) * nl]); sh.integrate(fallback_weights=True)
        did, anc = sh.resample(u01)
        gathered = [None] * world
        dist.all_gather_object(gathered, (sh.poses, sh.covs, sh.weights))
        if rank == 0:
            one.motion(1, u, 1.0, par)
            one.set_scan(r, ang); one.scan_match(); one.weight(z); one.integrate(fallback_weights=True)
            did1, anc1 = one.resample(u01)
            poses = np.concatenate([g[0] for g in gathered])
            covs = np.concatenate([g[1] for g in gathered])
            wts = np.concatenate([g[2] for g in gathered])
            good = (did == did1 and np.array_equal(anc, anc1) and np.array_equal(poses, one.poses)
                    and np.array_equal(covs, one.covs) and np.array_equal(wts, one.weights))
            ok &= good
            print("scan %2d resampled=%s identical=%s" % (step, did, good), flush=True)
        migrated += sh.migrated_particles
    # maps: every rank exports its particles' tiles, rank 0 compares with the single set
    mine = {}
    for j in range(nl):
        mine[rank * nl + j] = {c: sh.export_tile(j, *c) for c in sh.list_tiles(j)}
    allmaps = [None] * world
    dist.all_gather_object(allmaps, mine)
    tot_mig = torch.tensor([sh.migrated_particles], device="cuda")
    dist.all_reduce(tot_mig)
    if rank == 0:
        for part in allmaps:
            for j, tiles in part.items():
                ok &= sorted(tiles.keys()) == sorted(one.list_tiles(j))
                for c, a in tiles.items():
                    ok &= bool(np.array_equal(a, one.export_tile(j, *c)))
        print("transport %s: particles migrated: %d; sharded run identical to the single set: %s" % (sh.transport, int(tot_mig), ok), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
