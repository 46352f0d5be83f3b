#!/usr/bin/env python
"""Per-kernel timings at BASELINE shapes (CUDA events, rotating input sets larger than L2).
usage: python tools/bench_kernels.py [c2|c3|c4] [--only k1a,k1b,k2,k3,k3ex]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops

SHAPES = {"c2": (1, 32, 4096, 4096), "c3": (1, 32, 32768, 4096), "c4": (1, 40, 16384, 5120)}
name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c2"
only = None
for a in sys.argv:
    if a.startswith("--only"):
        only = set(sys.argv[sys.argv.index(a) + 1].split(","))
B, H, S, HD = SHAPES[name]
P, D = 128, 128
dev = torch.device("cuda")
PEAK = 6545.3


def timeit(fn, nset, iters=30, warm=5):
    for i in range(warm):
        fn(i % nset)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i % nset)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3   # us


def graph_time(fn, nset, iters=30):
    """same, but the loop is one CUDA graph (no Python launch overhead between kernels)"""
    for i in range(3):
        fn(i % nset)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % nset)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


res = {"shape": name, "B": B, "H": H, "S": S, "HD": HD}
gen = torch.Generator(device="cuda").manual_seed(0)
cfg = rtkv.CompressionConfig(num_hidden_layers=max(32, H), num_attention_heads=H, hidden_size=HD,
                             high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2)

if only is None or "k1a" in only:
    a_bytes = B * H * S * P * 4
    nset = max(2, int(400e6 // a_bytes) + 1)
    As = [torch.rand(B, H, S, P, generator=gen, device=dev) / P for _ in range(nset)]
    t = graph_time(lambda i: ops.prompt_mass(As[i], prompt_len=P), nset)
    res["k1a_us"] = t; res["k1a_GBs"] = a_bytes / t / 1e3; res["k1a_frac"] = res["k1a_GBs"] / PEAK
    del As

if only is None or "k1b" in only:
    q_bytes = B * H * S * D * 2
    nset = max(2, int(400e6 // q_bytes) + 1)
    Qs = [torch.randn(B, H, S, D, generator=gen, device=dev).bfloat16() for _ in range(nset)]
    Kp = torch.randn(B, H, P, D, generator=gen, device=dev).bfloat16()
    lse = torch.full((B, H, S), 8.0, device=dev)
    t = graph_time(lambda i: ops.prompt_mass_qk_partials(Qs[i], Kp, lse, D ** -0.5, True), nset)
    res["k1b_us"] = t; res["k1b_GBs"] = q_bytes / t / 1e3; res["k1b_frac_hbm"] = res["k1b_GBs"] / PEAK
    res["k1b_TFLOPs"] = 2.0 * B * H * S * P * D / t / 1e6
    if "--qk-bisect" in sys.argv:
        from rtkv import _lib
        for mode, tag in ((1, "no_exp"), (2, "no_tmem_ld"), (3, "no_mma"), (4, "tma_only")):
            _lib.lib.rtkv_debug_qk_mode(mode)
            res[f"k1b_{tag}_us"] = graph_time(lambda i: ops.prompt_mass_qk_partials(Qs[i], Kp, lse, D ** -0.5, True), nset)
        _lib.lib.rtkv_debug_qk_mode(0)
    del Qs

if only is None or "k2" in only:
    i = torch.arange(S, device=dev, dtype=torch.float32)
    mass = (torch.clamp(128.0 / (i + 1), max=1.0) * (1 + 0.05 * torch.rand(S, generator=gen, device=dev)))[None].repeat(B, 1).contiguous()
    for layer, tag in ((0, "fit"), (31, "walk")):
        c = rtkv.CompressionConfig(num_hidden_layers=32, high_precision_bits=16 if tag == "walk" else 8,
                                   medium_precision_bits=8 if tag == "walk" else 4, low_precision_bits=4 if tag == "walk" else 2)
        comp = rtkv.RealTimePrefillCompressor(c)
        params, _ = comp.layer_params(layer, S, ops.F32)
        sel = ops.score_select(params, B, S, dev, mass=mass)
        kept = int(sel.read_stats()[1]["n_selected"].sum())
        t = graph_time(lambda i: ops.score_select(params, B, S, dev, mass=mass), 1)
        res[f"k2_{tag}_us"] = t; res[f"k2_{tag}_kept"] = kept

if only is None or "k3" in only:
    kv_bytes = B * S * HD * 4
    nset = max(2, int(300e6 // (2 * kv_bytes)) + 1)
    Ks = [torch.randn(B, S, HD, generator=gen, device=dev) for _ in range(nset)]
    Vs = [torch.randn(B, S, HD, generator=gen, device=dev) for _ in range(nset)]
    ok, ov = torch.empty_like(Ks[0]), torch.empty_like(Ks[0])
    labels = torch.randint(0, 3, (B, S), generator=gen, device=dev, dtype=torch.uint8)
    t = graph_time(lambda i: ops.quant_gather(Ks[i], Vs[i], None, labels, (2, 4, 8), S, out0=ok, out1=ov), nset)
    res["k3_us"] = t; res["k3_GBs"] = 4 * kv_bytes / t / 1e3; res["k3_frac"] = res["k3_GBs"] / PEAK
    t = graph_time(lambda i: (ok.copy_(Ks[i]), ov.copy_(Vs[i])), nset)
    res["torch_copy_same_bytes_us"] = t; res["torch_copy_GBs"] = 4 * kv_bytes / t / 1e3
    for hd, tag in ((torch.bfloat16, "bf16"), (torch.float16, "f16")):
        Kh = [k.to(hd) for k in Ks] + [k.to(hd) * 0.5 for k in Ks]
        Vh = [v.to(hd) for v in Vs] + [v.to(hd) * 0.5 for v in Vs]
        okh, ovh = torch.empty_like(Kh[0]), torch.empty_like(Kh[0])
        t = graph_time(lambda i: ops.quant_gather(Kh[i], Vh[i], None, labels, (2, 4, 8), S, out0=okh, out1=ovh), len(Kh))
        res[f"k3_{tag}_us"] = t; res[f"k3_{tag}_frac"] = 2 * kv_bytes / t / 1e3 / PEAK
        if tag == "bf16":      # A/B: element-wise bf16 code (the round-1 kernel)
            ops.lib.rtkv_debug_bf16_packed(0)
            t = graph_time(lambda i: ops.quant_gather(Kh[i], Vh[i], None, labels, (2, 4, 8), S, out0=okh, out1=ovh), len(Kh))
            res["k3_bf16_elementwise_us"] = t; res["k3_bf16_elementwise_frac"] = 2 * kv_bytes / t / 1e3 / PEAK
            ops.lib.rtkv_debug_bf16_packed(1)
        del Kh, Vh

if only is None or "k3ex" in only:
    # extended scale granularities / outlier split (quant_ex.cu): same algorithmic bytes as K3
    kv_bytes = B * S * HD * 4
    nset = max(2, int(300e6 // (2 * kv_bytes)) + 1)
    Ks = [torch.randn(B, S, HD, generator=gen, device=dev) for _ in range(nset)]
    Vs = [torch.randn(B, S, HD, generator=gen, device=dev) for _ in range(nset)]
    labels = torch.randint(0, 3, (B, S), generator=gen, device=dev, dtype=torch.uint8)
    variants = {
        "group128": ops.quant_options("token_group", "token_group", group_size=128),
        "group128_tau6": ops.quant_options("token_group", "token_group", group_size=128, outlier_tau=6.0),
        "row_tau6": ops.quant_options("token", "token", outlier_tau=6.0),
        "chanK32_groupV128": ops.quant_options("channel", "token_group", group_size=128, token_group=32),
        "chanK16_chanV16": ops.quant_options("channel", "channel", token_group=16),
    }
    ok, ov = torch.empty_like(Ks[0]), torch.empty_like(Ks[0])
    from rtkv._lib import lib as _l
    _l.rtkv_debug_channel_variant(1)
    t = graph_time(lambda i: ops.quant_gather_ex(Ks[i], Vs[i], None, labels, (2, 4, 8), S, ops.quant_options("channel", "channel", token_group=32), out0=ok, out1=ov), nset)
    res["k3ex_chan32_both_staged_us"] = t
    _l.rtkv_debug_channel_variant(0)
    t = graph_time(lambda i: ops.quant_gather_ex(Ks[i], Vs[i], None, labels, (2, 4, 8), S, ops.quant_options("channel", "channel", token_group=32), out0=ok, out1=ov), nset)
    res["k3ex_chan32_both_l2_us"] = t
    for tag, opt in variants.items():
        t = graph_time(lambda i: ops.quant_gather_ex(Ks[i], Vs[i], None, labels, (2, 4, 8), S, opt, out0=ok, out1=ov), nset)
        res[f"k3ex_{tag}_us"] = t; res[f"k3ex_{tag}_frac"] = 4 * kv_bytes / t / 1e3 / PEAK
    # packed rows: read K fp32, write codes (mean 4.67 bits) + parameters
    for tag, opt in (("row", ops.quant_options()), ("group128", variants["group128"]),
                     ("chan32", ops.quant_options("channel", token_group=32))):
        t = graph_time(lambda i: ops.quant_pack_ex(Ks[i], None, labels, (2, 4, 8), S, opt), nset)
        res[f"pack_{tag}_us"] = t; res[f"pack_{tag}_read_GBs"] = kv_bytes / t / 1e3
    t = graph_time(lambda i: ops.quant_pack(Ks[i], None, labels, (2, 4, 8), S), nset)
    res["pack_reference_mode_us"] = t; res["pack_reference_mode_read_GBs"] = kv_bytes / t / 1e3

print(json.dumps(res))
