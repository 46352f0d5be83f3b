#!/usr/bin/env python
"""Stage timings of the one-exchange head-sharded layer on ONE GPU: rank 0's work for world = 2, 4, 8 with the other
ranks' records produced locally (LocalExchange), S = 32768, 7B shapes.  Prints JSON per world size."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops, dist as rd

dev = torch.device("cuda", 0)
B, H, S, D, P = 1, 32, int(os.environ.get("S", 32768)), 128, 128
bits = tuple(int(x) for x in os.environ.get("BITS", "2,4,8").split(","))
cfg = rtkv.CompressionConfig(high_precision_bits=bits[2], medium_precision_bits=bits[1], low_precision_bits=bits[0])
comp = rtkv.RealTimePrefillCompressor(cfg)
params, _ = comp.layer_params(5, S, ops.F32)
NBUF = 3   # rotate inputs so that nothing is served from L2


def timeit(fn, iters=10, pre=None, reps=6):
    """Median time of one call of fn, measured inside a CUDA graph of `reps` calls (no Python launch gaps);
    pre(it) runs (untimed, eagerly) before every replay."""
    for it in range(2):
        if pre is not None:
            pre(it)
        fn(it)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for it in range(reps):
            fn(it)
    ts = []
    for it in range(iters + 2):
        if pre is not None:
            pre(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


for R in [int(x) for x in os.environ.get("WORLDS", "2,4,8").split(",")]:
    hl, w = H // R, H * D // R
    g = torch.Generator(device="cuda").manual_seed(1)
    As = [torch.rand(B, hl, S, P, generator=g, device=dev) / P for _ in range(NBUF)]
    Ks = [torch.randn(B, S, w, generator=g, device=dev) for _ in range(NBUF)]
    Vs = [torch.randn(B, S, w, generator=g, device=dev) for _ in range(NBUF)]
    ok, ov = torch.empty_like(Ks[0]), torch.empty_like(Vs[0])
    xch = rd.LocalExchange(B, S, R, dev)
    views = [xch.view(r) for r in range(R)]
    gs = H // 8
    res = {"world": R, "S": S}

    def publish_others(it):
        for r in range(1, R):
            v = views[r]
            ops.shard_stats(As[it % NBUF], Ks[it % NBUF], Vs[it % NBUF], P, gs, v.ptrs, r, R, v.state)

    solo = rd.LocalExchange(B, S, R, dev).view(0)      # own epoch counter: the lock-step exchange below stays in step
    v0 = solo
    res["shard_stats_us"] = timeit(lambda it: ops.shard_stats(As[it % NBUF], Ks[it % NBUF], Vs[it % NBUF], P, gs, v0.ptrs, 0, R, v0.state))
    ops.lib.rtkv_debug_shard_variant(1)
    res["shard_stats_generic_us"] = timeit(lambda it: ops.shard_stats(As[it % NBUF], Ks[it % NBUF], Vs[it % NBUF], P, gs, v0.ptrs, 0, R, v0.state))
    ops.lib.rtkv_debug_shard_variant(0)
    v0 = views[0]
    byts = As[0].numel() * 4 + 2 * Ks[0].numel() * 4
    res["shard_stats_GBs"] = byts / res["shard_stats_us"] / 1e3
    # plain K1a partials of the same A slice, and a plain copy of K for reference
    res["k1a_partials_us"] = timeit(lambda it: ops.prompt_mass_partials(As[it % NBUF], P, gs))
    res["k1a_partials_GBs"] = As[0].numel() * 4 / res["k1a_partials_us"] / 1e3
    res["row_minmax_us"] = timeit(lambda it: ops.row_minmax(Ks[it % NBUF], Vs[it % NBUF], None, S))
    res["row_minmax_GBs"] = 2 * Ks[0].numel() * 4 / res["row_minmax_us"] / 1e3

    def pub_all(it):
        for r in range(R):
            v = views[r]
            ops.shard_stats(As[it % NBUF], Ks[it % NBUF], Vs[it % NBUF], P, gs, v.ptrs, r, R, v.state)

    def pub_and_combine(it):
        pub_all(it)
        ops.shard_combine(v0.local_ptr, v0.state, B, S, H, R, v0.mass, v0.minmax)
    t_pub = timeit(pub_all, reps=2)
    res["combine_us"] = timeit(pub_and_combine, reps=2) - t_pub
    mass, mm = v0.mass, v0.minmax
    sel = ops.score_select(params, B, S, dev, mass=mass)
    res["k2_us"] = timeit(lambda it: ops.score_select(params, B, S, dev, mass=mass))
    res["kept"] = int(sel.read_stats()[1]["n_selected"].sum())
    res["k3_token_minmax_us"] = timeit(lambda it: ops.quant_gather_token_minmax(Ks[it % NBUF], Vs[it % NBUF], sel, sel.labels, bits, S, mm,
                                                                              zero_pad=False, out0=ok, out1=ov))
    res["k3_GBs"] = res["kept"] * w * 4 * 4 / res["k3_token_minmax_us"] / 1e3
    res["k3_plain_us"] = timeit(lambda it: ops.quant_gather(Ks[it % NBUF], Vs[it % NBUF], sel, sel.labels, bits, S, zero_pad=False, out0=ok, out1=ov))
    res["sum_us"] = res["shard_stats_us"] + res["combine_us"] + res["k2_us"] + res["k3_token_minmax_us"]
    # rank 0's whole layer as it runs in a real job: the other ranks' records are already there (published first),
    # then stats -> combine -> K2 -> K3 back to back; external events inside the graph bracket rank 0's chain
    NL = 4
    ev = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)) for _ in range(NL)]

    # the other ranks' kernels read their OWN tensors (a second set), so that rank 0's L2 behaviour is not flattered
    A2, K2, V2 = torch.rand_like(As[0]), torch.randn_like(Ks[0]), torch.randn_like(Vs[0])

    def chain(timed):
        for l in range(NL):
            for r in range(1, R):
                v = views[r]
                ops.shard_stats(A2, K2, V2, P, gs, v.ptrs, r, R, v.state)
            if timed:
                ev[l][0].record()
            ops.shard_stats(As[l % NBUF], Ks[l % NBUF], Vs[l % NBUF], P, gs, v0.ptrs, 0, R, v0.state)
            m_, mm_ = ops.shard_combine(v0.local_ptr, v0.state, B, S, H, R, v0.mass, v0.minmax)
            sel_ = ops.score_select(params, B, S, dev, mass=m_)
            ops.quant_gather_token_minmax(Ks[l % NBUF], Vs[l % NBUF], sel_, sel_.labels, bits, S, mm_, zero_pad=False, out0=ok, out1=ov)
            if timed:
                ev[l][1].record()
    chain(False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain(True)
    acc = []
    for _ in range(6):
        g.replay()
        torch.cuda.synchronize()
        acc.append(sum(a.elapsed_time(b) for a, b in ev) * 1e3 / NL)
    acc.sort()
    res["chain_us_per_layer"] = acc[len(acc) // 2]
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    del As, Ks, Vs, xch, views
    torch.cuda.empty_cache()
