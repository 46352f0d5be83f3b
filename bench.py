#!/usr/bin/env python
"""bench.py — prefill KV-compress throughput on B200 (BASELINE.json metric).

A step = one pass of the hot path (K1 prompt_mass -> K2 score_select -> K3
quant_gather) over every decoder layer of one prefill.  Default workload is
BASELINE.json configs[1]: Llama-2-7B random-init, 32 layers x 4096 tokens, batch 1,
fp32 K/V (the reference runs the model in fp32), bits 8/4/2, theta 0.7/0.3.
N>1 (torchrun): every rank compresses its own sequence (batch sharding, no
collective on the data path) -> weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c4]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))

WORKLOADS = {
    "tiny": ("tiny self-test workload (not a BASELINE config)", 3, 256, 4, 512, 320, 1, (2, 4, 8)),
    # name: (description, layers, hidden, heads, intermediate, S, B, bits(low,mid,high))
    "c1": ("single layer K/V [1,32,2048,128] fp32 (configs[0])", 1, 4096, 32, 11008, 2048, 1, (2, 4, 8)),
    "c2": ("Llama-2-7B random-init, 32-layer prefill of 4096 tokens, batch 1 (configs[1])", 32, 4096, 32, 11008, 4096, 1, (2, 4, 8)),
    "c3": ("Llama-2-7B random-init, 32-layer prefill of 32768 tokens, batch 1, KV-head-sharded over the GPUs (configs[2])", 32, 4096, 32, 11008, 32768, 1, (2, 4, 8)),
    "c5": ("Llama-2-7B random-init, batch 8 per GPU x 4096-token prefill, batch-sharded (configs[4]: 64 sequences on 8 GPUs)", 32, 4096, 32, 11008, 4096, 8, (2, 4, 8)),
    "c4": ("Llama-2-13B random-init, 40-layer prefill of 16384 tokens, batch 1 (configs[3], mixed 8/4/2)", 40, 5120, 40, 13824, 16384, 1, (2, 4, 8)),
}
PEAKS_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return PEAKS_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- synthetic prefill
def rope_tables(S, D, device):
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, device=device, dtype=torch.float32) / D))
    ang = torch.arange(S, device=device, dtype=torch.float32)[:, None] * inv[None, :]
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def rope(x, cos, sin):
    h = x.shape[-1] // 2
    rot = torch.cat([-x[..., h:], x[..., :h]], dim=-1)
    return x * cos + rot * sin


def rmsnorm(x):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)


@torch.no_grad()
def llama_prefill_inputs(n_layers, hidden, heads, inter, S, B, P, seed, device, keep_layers=None, head_slice=None):
    """Random-init Llama-2 stack (N(0,0.02) weights, RoPE, causal softmax attention, SwiGLU MLP)
    run in plain torch fp32 as SETUP: yields per layer the K/V the compressor sees
    ([B,S,H*D] after RoPE, as modified_llama.py:105-108) and the first P key columns of the
    softmaxed attention ([B,H,S,P]; the reference never reads any other column)."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = hidden // heads
    ids = torch.randint(0, 32000, (B, S), generator=g, device=device)
    emb = torch.randn(32000, hidden, generator=g, device=device) * 0.02
    x = emb[ids]
    del emb
    cos, sin = rope_tables(S, D, device)
    out = []
    rows = max(128, min(S, (1 << 28) // (heads * S)))       # query rows per softmax block (<= 1 GiB logits)
    for l in range(n_layers):
        def w(o, i):
            return torch.randn(o, i, generator=g, device=device) * 0.02
        h = rmsnorm(x)
        q = (h @ w(hidden, hidden).T).view(B, S, heads, D).transpose(1, 2)
        k = (h @ w(hidden, hidden).T).view(B, S, heads, D).transpose(1, 2)
        v = (h @ w(hidden, hidden).T).view(B, S, heads, D).transpose(1, 2)
        q, k = rope(q, cos, sin), rope(k, cos, sin)
        A = torch.empty(B, heads, S, P, device=device)
        ctx = torch.empty(B, heads, S, D, device=device)
        for r0 in range(0, S, rows):
            r1 = min(S, r0 + rows)
            logits = q[:, :, r0:r1] @ k.transpose(-1, -2) / math.sqrt(D)
            causal = torch.arange(S, device=device)[None, :] > torch.arange(r0, r1, device=device)[:, None]
            logits.masked_fill_(causal, float("-inf"))
            probs = torch.softmax(logits, dim=-1)
            A[:, :, r0:r1] = probs[..., :P]
            ctx[:, :, r0:r1] = probs @ v
            del logits, probs
        if keep_layers is None or l in keep_layers:
            if head_slice is None:
                out.append((l, k.transpose(1, 2).reshape(B, S, hidden).contiguous(),
                            v.transpose(1, 2).reshape(B, S, hidden).contiguous(), A))
            else:   # this rank's contiguous slice of the heads (identical data on every rank)
                hr, hw = head_slice
                hl = heads // hw
                out.append((l, k[:, hr * hl:(hr + 1) * hl].transpose(1, 2).reshape(B, S, hl * D).contiguous(),
                            v[:, hr * hl:(hr + 1) * hl].transpose(1, 2).reshape(B, S, hl * D).contiguous(),
                            A[:, hr * hl:(hr + 1) * hl].contiguous()))
        x = x + ctx.transpose(1, 2).reshape(B, S, hidden) @ w(hidden, hidden).T
        h = rmsnorm(x)
        gate, up = h @ w(inter, hidden).T, h @ w(inter, hidden).T
        x = x + (torch.nn.functional.silu(gate) * up) @ w(hidden, inter).T
        del q, k, v, ctx, gate, up, h
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_baseline(layers_host, cfg, bits, S, n_layers_total, B, reps=1):
    """Oracle port of the reference's CPU path (quantise every token, then select and gather, as
    the reference does) timed on the host cores over a bounded sample of layers."""
    import oracle
    from rtkv.selective_propagation import SelectiveTokenPropagator
    oracle.set_num_threads(os.cpu_count() or 1)
    prop = SelectiveTokenPropagator(cfg)
    P = max(1, min(S // 5, 128))
    l0, K0, V0, A0 = layers_host[0]
    oracle.compress_layer(K0[:, :256], V0[:, :256], oracle.F32, A0[:, :, :256], oracle.F32, min(P, 51),
                          cfg.alpha, cfg.beta, cfg.gamma, 1.0, cfg.theta_h, cfg.theta_m, bits, 0.8)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        for (l, K, V, A) in layers_host:
            oracle.compress_layer(K, V, oracle.F32, A, oracle.F32, P, cfg.alpha, cfg.beta, cfg.gamma,
                                  cfg.layer_weights[l], cfg.theta_h, cfg.theta_m, bits,
                                  prop.get_layer_propagation_ratio(l))
    dt = (time.perf_counter() - t0) / reps
    per_layer = dt / len(layers_host)
    return B * S / (per_layer * n_layers_total), per_layer, oracle.num_threads()


# ----------------------------------------------------------------------------- head-sharded mode
def bench_head_sharded(args, layers, comp, cfg, config, dev, rank, world, L, S, B, P, hidden, heads, bits):
    """BASELINE configs[2]: one long sequence, KV heads split over the ranks (strong scaling).
    Per layer: K1a partials -> all-gather -> K2 (replicated, identical on every rank) ->
    row min/max of the kept rows -> MIN all-reduce -> K3.  No host synchronisation."""
    import torch.distributed as dist
    from rtkv import dist as rd
    p2p = args.p2p
    sharded = rd.HeadShardedCompressor(comp, heads, p2p=p2p)
    if p2p:
        try:
            sharded._mailbox(B, S, dev)
        except Exception as e:            # symmetric memory unavailable: fall back to the collective scheme
            sys.stderr.write(f"peer mailbox unavailable ({e}); using the all-reduce scheme\n")
            sharded = rd.HeadShardedCompressor(comp, heads, p2p=False)
            p2p = False
    w = hidden // world
    outs = [(torch.empty(B, S, w, device=dev), torch.empty(B, S, w, device=dev)) for _ in range(2)]

    def step():
        sels = []
        for j, (l, K, V, A) in enumerate(layers):
            sels.append(sharded.compress_layer_async(K, V, A, l, outs[j & 1][0], outs[j & 1][1]))
        return sels

    for _ in range(args.warmup):
        sels = step()
    torch.cuda.synchronize()
    mode = "eager launches"
    graph = None
    try:                                   # NCCL collectives can be captured; fall back to eager if not
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        graph.replay()
        torch.cuda.synchronize()
        mode = "one CUDA graph per step (kernels + NCCL collectives)"
    except Exception as e:                 # pragma: no cover
        sys.stderr.write(f"graph capture of the head-sharded step failed ({e}); timing eager launches\n")
        graph = None
        torch.cuda.synchronize()
    run = graph.replay if graph is not None else step
    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        run()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    kept = [int(s.read_stats()[1]["n_selected"].sum()) for s in sels]
    e = 4
    path_bytes = sum(n * 2 * hidden * 2 * e for n in kept) + L * (B * S * heads * P * e + B * S * 16)
    peak, peak_src = measured_hbm_peak()
    line = {"metric": "prefill_kv_compress_tokens_per_s", "value": B * S / (ms_per_step * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "clocks": clocks, "gpu_launches": (4 if p2p else 5) * L * args.steps, "launch_mode": mode,
            "roofline": {"bound": "hbm", "kernel": "whole path, all ranks", "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                         "peak_source": peak_src + f" x {world} GPUs", "traffic": None,
                         "retained_fraction": sum(kept) / (L * B * S)},
            "exchange": ("fused: row min/max pushed to peer mailboxes over NVLink inside the quantise kernel (K/V read once)"
                         if p2p else "two-pass: row_minmax kernel + MIN all-reduce + quantise (K/V read twice)"),
            "collectives_per_layer": {"all_gather_bytes": B * S * (8 // world) * 4,
                                      "all_reduce_bytes": 0 if p2p else B * S * 16,
                                      "p2p_mailbox_bytes_per_rank": B * S * 2 * 16 * (world - 1) if p2p else 0}}
    if rank == 0:
        _emit(line)
    # a live CUDA graph that holds NCCL kernels can dead-lock communicator teardown: drop it first
    del run, graph
    torch.cuda.synchronize()
    _shutdown(dist)


def _shutdown(dist):
    """Tear the process group down, but never hang the box: hard-exit if teardown stalls."""
    import threading
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        dist.destroy_process_group()
    finally:
        os._exit(0)


# ----------------------------------------------------------------------------- main
def _emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints there
    (e.g. "NCCL version ...") was redirected to stderr by _quiet_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="rtkv", choices=["rtkv", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=0, help="override the number of layers (setup time)")
    ap.add_argument("--bits", default="", help="low,mid,high bit widths, e.g. 8,8,8 (configs[3] precision sweep)")
    ap.add_argument("--ablation", default="", choices=["", "quant_only"],
                    help="quant_only: propagation ratios 1.0, every token kept (configs[4] ablation)")
    ap.add_argument("--p2p", action="store_true", help="head-sharded mode: fused peer-memory min/max exchange inside the quantise kernel "
                    "instead of the two-pass min/max + MIN all-reduce scheme")
    ap.add_argument("--prefetch-mb", type=int, default=0,
                    help="prefetch this many MB of K into L2 on a side stream during the selection kernel (measured: no gain)")
    ap.add_argument("--no-pdl", action="store_true", help="launch without programmatic dependent launch (A/B)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "rtkv" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, L, hidden, heads, inter, S, B, bits = WORKLOADS[args.workload]
    if args.layers > 0:
        L = args.layers
    if args.bits:
        bits = tuple(int(x) for x in args.bits.split(","))
    head_sharded = args.workload == "c3" and world > 1
    P = max(1, min(S // 5, 128))
    sample_layers = sorted({0, L // 3, (2 * L) // 3, L - 1})

    import rtkv
    ratios = {"early_layer_ratio": 1.0, "middle_layer_ratio": 1.0, "later_layer_ratio": 1.0} if args.ablation == "quant_only" else {}
    cfg = rtkv.CompressionConfig(num_hidden_layers=max(L, 2), num_attention_heads=heads, hidden_size=hidden,
                                 high_precision_bits=bits[2], medium_precision_bits=bits[1],
                                 low_precision_bits=bits[0], **ratios)
    config = {"workload": desc, "layers": L, "seq_len": S, "batch_per_gpu": B, "kv_width": hidden, "heads": heads,
              "prompt_len": P, "bits_low_mid_high": list(bits), "theta_h": cfg.theta_h, "theta_m": cfg.theta_m,
              "ratios": [cfg.early_layer_ratio, cfg.middle_layer_ratio, cfg.later_layer_ratio],
              "kv_dtype": "float32", "attention_input": f"[B,H,S,P={P}] probability slice, float32",
              "sharding": ("KV heads (2 small NCCL exchanges per layer: score partials, row min/max)" if head_sharded else
                           "batch (one sequence per GPU, no collective)" if world > 1 else "single GPU"),
              "l2": "per-step working set (inputs+outputs, all layers) far exceeds the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun pins OMP_NUM_THREADS=1; the reference arm uses every host core (rank 0 runs alone)
        import oracle
        oracle.set_num_threads(os.cpu_count() or 1)
        dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
        if dev.type == "cuda":
            torch.cuda.set_device(dev)
        layers = llama_prefill_inputs(L, hidden, heads, inter, S, B, P, 1234, dev, keep_layers=set(sample_layers))
        host = [(l, K.cpu().numpy(), V.cpu().numpy(), A.cpu().numpy()) for (l, K, V, A) in layers]
        del layers
        vals = []
        for it in range(args.warmup + args.steps):
            v, per_layer, cores = cpu_baseline(host, cfg, bits, S, L, B)
            if it >= args.warmup:
                vals.append((v, per_layer))
        value = float(np.mean([v for v, _ in vals]))
        per_layer = float(np.mean([p for _, p in vals]))
        line = {"impl": "reference", "metric": "prefill_kv_compress_tokens_per_s", "value": value, "unit": "tokens/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_layer * len(sample_layers) * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
                                 "sample": f"layers {sample_layers} of {L} per step, C/OpenMP port of the reference's "
                                           f"CPU path (oracle/), per-layer time x {L} layers"},
                "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    # ------------------------------------------------------------------ rtkv arm (GPU)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    layers = llama_prefill_inputs(L, hidden, heads, inter, S, B, P, 1234 + (0 if head_sharded else rank), dev,
                                  head_slice=(rank, world) if head_sharded else None)
    torch.cuda.synchronize()
    comp = rtkv.RealTimePrefillCompressor(cfg)
    if args.no_pdl:
        from rtkv import _lib as _rl
        _rl.lib.rtkv_debug_pdl(0)
    if head_sharded:
        return bench_head_sharded(args, layers, comp, cfg, config, dev, rank, world, L, S, B, P, hidden, heads, bits)
    outs = [(torch.empty_like(K), torch.empty_like(V)) for (_, K, V, _) in layers]
    ev_k3 = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in layers]
             for _ in range(args.steps)]
    from rtkv import ops

    def step(events=None):
        sels = []
        for j, (l, K, V, A) in enumerate(layers):
            params, _ = comp.layer_params(l, S, ops.F32)
            mass = ops.prompt_mass(A, prompt_len=P)
            sel = ops.score_select(params, B, S, dev, mass=mass)
            if events is not None:
                events[j][0].record()
            ops.quant_gather(K, V, sel, sel.labels, bits, S, zero_pad=False, out0=outs[j][0], out1=outs[j][1])
            if events is not None:
                events[j][1].record()
            sels.append(sel)
        return sels

    for _ in range(args.warmup):
        sels = step()
    torch.cuda.synchronize()

    # The timed region replays one CUDA graph of the whole step (3 kernels x L layers, one stream,
    # layer order preserved): launch-bound inner loops belong in graphs, not in Python.
    def capture(two_stream, events=None):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev) if two_stream else None
        pf = torch.cuda.Stream(device=dev) if args.prefetch_mb > 0 else None
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            out = []
            for j, (l, K, V, A) in enumerate(layers):
                params, _ = comp.layer_params(l, S, ops.F32)
                mass = ops.prompt_mass(A, prompt_len=P)
                if pf is not None:      # L2 warm-up of the first K rows while K2 (one SM) runs
                    pf.wait_stream(main)
                    with torch.cuda.stream(pf):
                        ops.prefetch_l2(K, args.prefetch_mb << 20)
                sel = ops.score_select(params, B, S, dev, mass=mass)
                if pf is not None:
                    main.wait_stream(pf)
                if side is None:
                    if events is not None:
                        events[j][0].record()
                    ops.quant_gather(K, V, sel, sel.labels, bits, S, zero_pad=False, out0=outs[j][0], out1=outs[j][1])
                    if events is not None:
                        events[j][1].record()
                else:  # K3 of layer j overlaps K1/K2 of layer j+1
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        ops.quant_gather(K, V, sel, sel.labels, bits, S, zero_pad=False, out0=outs[j][0], out1=outs[j][1])
                out.append(sel)
            if side is not None:
                main.wait_stream(side)
        return g, out

    def time_graph(g):
        for _ in range(args.warmup):
            g.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / args.steps

    graph, sels = capture(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step = time_graph(graph)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * S / (ms_per_step * 1e-3)
    graph2, _ = capture(True)
    ms_two_stream = time_graph(graph2)

    # K3 launch durations inside the graph: event-record nodes around every K3 kernel
    k3_graph_ms = None
    try:
        gev = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
               for _ in layers]
        graph3, _ = capture(False, gev)
        for _ in range(3):
            graph3.replay()
        acc = []
        for _ in range(args.steps):
            graph3.replay()
            torch.cuda.synchronize()
            acc.append(sum(a.elapsed_time(b) for a, b in gev))
        k3_graph_ms = float(np.mean(acc))
        del graph3
    except Exception as e:     # pragma: no cover  (timing events in graphs not supported: keep the eager figure)
        sys.stderr.write(f"event nodes in the graph failed ({e}); K3 is timed in the eager pass only\n")

    # instrumented eager pass (same process, same data): CUDA events around every K3 launch
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for it in range(args.steps):
        sels = step(ev_k3[it])
    t_end.record()
    torch.cuda.synchronize()
    eager_ms_per_step = t_start.elapsed_time(t_end) / args.steps

    # measured retained rows -> algorithmic bytes (SURVEY.md §8d formula, measured rho)
    kept = [int(s.read_stats()[1]["n_selected"].sum()) for s in sels]
    e = 4
    k3_bytes = [n * 2 * hidden * (e + e) for n in kept]
    k1_bytes = B * S * heads * P * e + B * S * 4
    k2_bytes = B * S * 12
    path_bytes = sum(k3_bytes) + L * (k1_bytes + k2_bytes)
    k3_eager_ms = float(np.mean([sum(a.elapsed_time(b) for a, b in evs) for evs in ev_k3]))
    k3_ms = k3_graph_ms if k3_graph_ms is not None else k3_eager_ms
    peak, peak_src = measured_hbm_peak()
    achieved = sum(k3_bytes) / (k3_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "quant_gather_vec_kernel (K3)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                "launch_ms_avg": k3_ms / L, "bytes_per_launch_avg": sum(k3_bytes) / L,
                "k3_share_of_step": k3_ms / (ms_per_step if k3_graph_ms is not None else eager_ms_per_step),
                "measured": ("CUDA event-record nodes around every K3 kernel inside the step graph" if k3_graph_ms is not None
                             else "CUDA events around every K3 launch in an eager pass of the same run"),
                "eager_pass": {"k3_ms_per_step": k3_eager_ms, "achieved": sum(k3_bytes) / (k3_eager_ms * 1e-3) / 1e9,
                               "note": "same kernels launched from Python; includes CPU launch gaps"},
                "retained_fraction": sum(kept) / (L * B * S),
                "whole_path": {"achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "frac_of_8TBs_nominal": path_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                               "bytes_per_step": path_bytes}}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("quant_gather_bytes_per_launch")
    except Exception:
        pass

    line = {"metric": "prefill_kv_compress_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": 3 * L * args.steps, "roofline": roofline,
            "launch_mode": "one CUDA graph per step (3 kernels x %d layers, single stream%s)" % (
                L, ("" if args.no_pdl else ", programmatic dependent launch") +
                ("" if args.prefetch_mb <= 0 else ", L2 prefetch of %d MB of K on a side stream during K2" % args.prefetch_mb)),
            "eager_ms_per_step": eager_ms_per_step,
            "two_stream": {"ms_per_step": ms_two_stream, "value": world * B * S / (ms_two_stream * 1e-3),
                           "note": "same graph with K3 of layer l on a second stream, overlapping K1/K2 of layer l+1"},
            "reference_recorded": {"tokens_per_s": 173, "note": "BASELINE.md: derived from the reference's recorded "
                                   "runs (23.7 s per 4096-token prefill, bits 16/8/4, unnamed GPU); not a published number"}}

    # ------------------------------------------------------------------ e2e: public API, host buffers
    if not args.no_e2e:
        host_in = [(l, K.cpu().pin_memory(), V.cpu().pin_memory(), A.cpu().pin_memory()) for (l, K, V, A) in layers]
        ids = torch.zeros(B, S, dtype=torch.long)
        ko = [torch.empty(B * S * hidden, dtype=torch.float32).pin_memory() for _ in range(2)]
        vo = [torch.empty(B * S * hidden, dtype=torch.float32).pin_memory() for _ in range(2)]
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            d2h = 0
            for j, (l, K, V, A) in enumerate(host_in):
                k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, l, out_keys=ko[j & 1], out_values=vo[j & 1])
                d2h += 2 * k_sel.numel() * 4 + B * S * 6 + 16 + 72 * B
            return d2h
        e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d2h = e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        h2d = sum(K.numel() * 4 + V.numel() * 4 + A.numel() * 4 for (_, K, V, A) in host_in)
        line["e2e"] = {"value": world * B * S / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                       "steps": e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "RealTimePrefillCompressor.compress_layer_kv_cache on pinned host tensors "
                              "(rtkv_compress_layer_host), synchronous per layer"}
        del host_in, ko, vo

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host = [(l, K.cpu().numpy(), V.cpu().numpy(), A.cpu().numpy()) for (l, K, V, A) in layers if l in sample_layers]
        v, per_layer, cores = cpu_baseline(host, cfg, bits, S, L, B)
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                                "sample": f"layers {sample_layers} of {L}, once each, C/OpenMP port of the reference's "
                                          f"CPU path (oracle/): {per_layer * 1e3:.1f} ms per layer x {L} layers"}
    if rank == 0:
        _emit(line)
    if world > 1:
        _shutdown(dist)


if __name__ == "__main__":
    main()
