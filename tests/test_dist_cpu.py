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
        # the whole head-sharded layer from the exchanged statistics: scores / labels / selection from the combined mass
        # (replicated), the rank's column slice quantised with the row-wide min/max, kept rows gathered -> equal, bit for bit, to
        # the same columns of the unsharded layer (token_importance.py:45 and dynamic_quantization.py:183-184 are the two
        # couplings across heads; everything else is per column slice)
        alpha, beta, gamma, w_l, th, tm, bits, ratio = 0.4, 0.3, 0.3, 0.75, 0.7, 0.3, (4, 8, 16), 0.6
        full = oracle.compress_layer(K, V, oracle.F32, A, oracle.F32, P, alpha, beta, gamma, w_l, th, tm, bits, ratio)
        mass = oracle.mass_from_partials(gathered, H)
        scores = oracle.importance_scores(mass, P, alpha, beta, gamma, w_l)
        labels, _ = oracle.assign_precision(scores, th, tm)
        mask, counts, _, _, _ = oracle.select_tokens(scores, labels, bits, S * ratio)
        ok3 = np.array_equal(scores, full["scores"]) and np.array_equal(labels, full["labels"]) and np.array_equal(mask, full["mask"])
        ok3 = ok3 and int(counts.min()) > 0 and 0 < int(counts.max()) < S       # the budget binds and no fallback is involved
        levels = (np.float32(2.0) ** np.asarray(bits, np.float32)[labels] - np.float32(1.0)).astype(np.float32)   # [B,S]

        def fake_quant(x, mn, mx):                           # dynamic_quantization.py:62-126, one fp32 rounding per operation
            flat = mx == mn
            scale = np.where(flat, np.float32(1.0), (mx - mn) / levels).astype(np.float32)
            zp = np.where(flat, np.float32(0.0), np.float32(0.0) - mn / scale).astype(np.float32)
            qv = np.clip(np.rint(x / scale[..., None] + zp[..., None]), np.float32(0.0), levels[..., None]).astype(np.float32)
            return ((qv - zp[..., None]) * scale[..., None]).astype(np.float32)

        ms = int(counts.max())
        for x, lo, hi, name in ((kl, mm[..., 0], -mm[..., 1], "k_sel"), (vl, mm[..., 2], -mm[..., 3], "v_sel")):
            fq = fake_quant(np.ascontiguousarray(x), lo, hi)
            got = oracle.gather_rows(fq, oracle.F32, mask, ms)
            ok3 = ok3 and np.array_equal(got, full[name][..., rank * w:(rank + 1) * w])
        # batch sharding covers every row exactly once
        sl = rd.batch_shard(5, rank, world)
        q.put((rank, bool(ok1), bool(ok2) and bool(ok3), (sl.start, sl.stop)))
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
