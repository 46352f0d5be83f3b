"""CPU, world_size 2, gloo: the exchange logic of rtkv/dist.py (the only multi-rank code) on host
tensors produced by the oracle; plus the shard arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rtkv import dist as rd
        H, S, P, B, D = 16, 96, 19, 2, 8
        rng = np.random.default_rng(7)                       # same data on every rank
        A = rng.random((B, H, S, P), dtype=np.float32)
        A /= A.sum(-1, keepdims=True)
        K = rng.standard_normal((B, S, H * D)).astype(np.float32)
        V = rng.standard_normal((B, S, H * D)).astype(np.float32)
        h0, h1, gs = rd.head_shard(H, rank, world)
        # exchange 1: per-group partial sums -> identical mass on every rank, equal to the unsharded oracle
        local = oracle.prompt_mass_partials(np.ascontiguousarray(A[:, h0:h1]), P, gs)
        gathered = rd.gather_partials(torch.from_numpy(local)).numpy()          # [R,B,S,8/R], rank-major
        gathered = np.ascontiguousarray(np.moveaxis(gathered, 0, -2)).reshape(B, S, 8)   # -> [B,S,8] in head order
        full_mass, full_parts = oracle.prompt_mass_groups(A, oracle.F32, P)
        ok1 = np.array_equal(gathered, full_parts) and np.array_equal(oracle.mass_from_partials(gathered, H), full_mass)
        # exchange 2: row min/max over the ranks' column slices == min/max of the whole row
        w = H * D // world
        kl, vl = K[..., rank * w:(rank + 1) * w], V[..., rank * w:(rank + 1) * w]
        mm_local = np.stack([kl.min(-1), -kl.max(-1), vl.min(-1), -vl.max(-1)], axis=-1).astype(np.float32)
        mm = rd.reduce_minmax(torch.from_numpy(mm_local)).numpy()
        want = np.stack([K.min(-1), -K.max(-1), V.min(-1), -V.max(-1)], axis=-1)
        ok2 = np.array_equal(mm, want)
        # batch sharding covers every row exactly once
        sl = rd.batch_shard(5, rank, world)
        q.put((rank, bool(ok1), bool(ok2), (sl.start, sl.stop)))
    finally:
        dist.destroy_process_group()


def test_head_shard_exchange_world2_gloo(built_lib):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True] and [r[2] for r in res] == [True, True]
    assert [r[3] for r in res] == [(0, 3), (3, 5)]


def test_shard_arithmetic(built_lib):
    from rtkv import dist as rd
    assert rd.head_shard(32, 3, 4) == (24, 32, 4)
    assert rd.head_shard(40, 0, 8) == (0, 5, 5)
    with pytest.raises(ValueError):
        rd.head_shard(12, 0, 2)
    with pytest.raises(ValueError):
        rd.head_shard(32, 0, 3)
    rows = [rd.batch_shard(64, r, 8) for r in range(8)]
    assert [(s.start, s.stop) for s in rows] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert rd.batch_shard(3, 3, 4) == slice(3, 3)
    # single process: the exchanges are identities
    t = torch.arange(8.0).view(1, 2, 4)
    assert torch.equal(rd.gather_partials(t), t.unsqueeze(0)) and torch.equal(rd.reduce_minmax(t), t)
