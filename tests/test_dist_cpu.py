"""CPU (gloo, world_size 2): the host-side logic of the sharded resample --
every rank derives the same migration plan from the all-gathered weights, and
executing the plan reproduces the global resample."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from thesis_b200.dist import plan_migration


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    w_local = torch.from_numpy(rng.normal(0, 1, n_local) * 1e5)
    ids_local = torch.arange(rank * n_local, (rank + 1) * n_local, dtype=torch.int64)   # particle "state"
    w_all = torch.empty(world * n_local, dtype=torch.float64)
    dist.all_gather_into_tensor(w_all, w_local)
    rc, anc = O.resample(w_all.numpy(), 0.4321)            # identical on every rank (the GPU runs resample_plan_kernel)
    send, recv = plan_migration(anc, n_local, rank, world)
    new = torch.full((n_local,), -1, dtype=torch.int64)
    # local ancestors
    for j in range(n_local):
        a = anc[rank * n_local + j] - rank * n_local
        if 0 <= a < n_local:
            new[j] = ids_local[a]
    # exchange: each peer receives the states it asked for
    for r in range(world):
        if r == rank:
            continue
        out = ids_local[torch.from_numpy(send[r].astype(np.int64))]
        dst_slots, rec_idx, n_in = recv[r]
        inc = torch.empty(n_in, dtype=torch.int64)
        if rank < r:
            dist.send(out, r)
            dist.recv(inc, r)
        else:
            dist.recv(inc, r)
            dist.send(out, r)
        new[torch.from_numpy(dst_slots.astype(np.int64))] = inc[torch.from_numpy(rec_idx.astype(np.int64))]
    q.put((rank, anc.copy(), new.numpy().copy(), {r: len(v) for r, v in send.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_resample_plan_world2():
    world, n_local = 2, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    anc0, anc1 = res[0][1], res[1][1]
    assert np.array_equal(anc0, anc1)                       # same ancestors on all ranks
    assert np.all(np.diff(anc0) >= 0)
    got = np.concatenate([res[0][2], res[1][2]])
    assert np.array_equal(got, anc0)                        # slot j now holds particle anc[j]
    assert (got >= 0).all()


def test_plan_is_consistent_between_ranks():
    rng = np.random.default_rng(3)
    world, n_local = 4, 32
    anc = np.sort(rng.integers(0, world * n_local, world * n_local))
    plans = [plan_migration(anc, n_local, r, world) for r in range(world)]
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            send_ab = plans[a][0][b]
            dst, rec, n_in = plans[b][1][a]
            assert n_in == len(send_ab)
            # what b expects from a, resolved through a's send list, is the global ancestor
            assert np.array_equal(send_ab[rec] + a * n_local, anc[dst + b * n_local])
    # the pulling receiver's source list is the sender's send list (same slots, same order)
    for b in range(world):
        _, recv_b, src_b = plan_migration(anc, n_local, b, world, want_sources=True)
        for a in range(world):
            if a != b:
                assert np.array_equal(src_b[a], plans[a][0][b])
                assert recv_b[a][2] == len(src_b[a])
    # a sub-set with no cross-rank ancestors moves nothing
    ident = np.arange(world * n_local)
    s, r = plan_migration(ident, n_local, 1, world)
    assert all(len(v) == 0 for v in s.values()) and all(v[2] == 0 for v in r.values())
