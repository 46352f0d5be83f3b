#!/usr/bin/env python
"""bench.py — prefill KV-compress throughput on B200 (BASELINE.json metric).

Metric (BASELINE.json): prefill KV-compress tokens/s and HBM GB/s, **Llama-2-7B, 32k context**.
A step = one pass of the hot path over every decoder layer of one prefill.

  N = 1   the metric's own configuration: random-init Llama-2-7B, 32 layers x 32768 tokens, batch 1, fp32 K/V
          (the reference runs the model in fp32), bits 8/4/2, theta 0.7/0.3; K1a prompt_mass -> K2 score_select ->
          K3 quant_gather per layer, one CUDA graph per step.  Secondary keys: the reference-default widths 16/8/4
          (selection + gather bind), the fused Q.K scoring kernel (K1b, tcgen05), e2e from host buffers, the CPU
          baselines.
  N > 1   BASELINE configs[2]: the same 32k prefill with the KV heads sharded over the N GPUs (STRONG scaling: the
          problem is the same as at N = 1), one in-kernel exchange per layer over NVLink symmetric memory
          (rtkv_shard_stats / rtkv_shard_combine; no NCCL call on the data path).  Secondary key: batch sharding
          (one 4k sequence per GPU, configs[4]-style, no exchange at all).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|c5] ...

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
    "c3": ("Llama-2-7B random-init, 32-layer prefill of 32768 tokens (32k ctx, the metric's configuration), batch 1; "
           "KV-head-sharded when run on several GPUs (configs[2])", 32, 4096, 32, 11008, 32768, 1, (2, 4, 8)),
    "c5": ("Llama-2-7B random-init, batch 8 per GPU x 4096-token prefill, batch-sharded (configs[4]: 64 sequences on 8 GPUs)", 32, 4096, 32, 11008, 4096, 8, (2, 4, 8)),
    "c4": ("Llama-2-13B random-init, 40-layer prefill of 16384 tokens, batch 1 (configs[3], mixed 8/4/2)", 40, 5120, 40, 13824, 16384, 1, (2, 4, 8)),
}
PEAKS_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
METRIC = "prefill_kv_compress_tokens_per_s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1604.7)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return PEAKS_FALLBACK_GBS, 1600.0, "fallback (B200_PROFILING.md)"


def measured_bf16_sustained():
    """The sustained cuBLAS bf16 figure of MEASURED_PEAKS.json (back-to-back matmuls for seconds), or None."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            v = json.load(f).get("bf16_tflops_sustained")
        return float(v) if v else None
    except Exception:
        return None


def measured_hbm_peak():
    p, _, src = measured_peaks()
    return p, src


# ----------------------------------------------------------------------------- reference-shaped config (no rtkv import)
class PathConfig:
    """The reference's CompressionConfig defaults that the path reads (configs/base_config.py:4-57), restated so that
    the reference arm does not have to import the rtkv package (which loads librtkv.so)."""
    alpha, beta, gamma = 0.4, 0.3, 0.3
    theta_h, theta_m = 0.7, 0.3
    early_layer_ratio, middle_layer_ratio, later_layer_ratio = 0.8, 0.6, 0.4

    def __init__(self, num_layers, bits, ratios=None):
        self.num_hidden_layers = num_layers
        self.low_precision_bits, self.medium_precision_bits, self.high_precision_bits = bits
        self.layer_weights = [1.0 - 0.5 * i / (num_layers - 1) for i in range(num_layers)] if num_layers > 1 else [1.0]
        if ratios:
            self.early_layer_ratio, self.middle_layer_ratio, self.later_layer_ratio = ratios

    def ratio(self, l):
        """selective_propagation.py:20-38"""
        L = self.num_hidden_layers
        if l < int(0.3 * L):
            return self.early_layer_ratio
        return self.middle_layer_ratio if l < int(0.7 * L) else self.later_layer_ratio


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
def llama_prefill_inputs(n_layers, hidden, heads, inter, S, B, P, seed, device, keep_layers=None, head_slice=None,
                         qk_layer=None):
    """Random-init Llama-2 stack (N(0,0.02) weights, RoPE, causal softmax attention, SwiGLU MLP)
    run in plain torch fp32 as SETUP: yields per layer the K/V the compressor sees
    ([B,S,H*D] after RoPE, as modified_llama.py:105-108) and the first P key columns of the
    softmaxed attention ([B,H,S,P]; the reference never reads any other column).
    qk_layer: additionally return that layer's (Q bf16, K_prompt bf16, LSE f32) for the fused-scoring kernel."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = hidden // heads
    ids = torch.randint(0, 32000, (B, S), generator=g, device=device)
    emb = torch.randn(32000, hidden, generator=g, device=device) * 0.02
    x = emb[ids]
    del emb
    cos, sin = rope_tables(S, D, device)
    out, qk = [], None
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
        want_qk = qk_layer is not None and l == qk_layer
        if want_qk:       # the fused-scoring kernel works on the half-precision operands a bf16 model would hold
            qb, kb = q.to(torch.bfloat16), k.to(torch.bfloat16)
            lse = torch.empty(B, heads, S, device=device)
        for r0 in range(0, S, rows):
            r1 = min(S, r0 + rows)
            causal = torch.arange(S, device=device)[None, :] > torch.arange(r0, r1, device=device)[:, None]
            logits = q[:, :, r0:r1] @ k.transpose(-1, -2) / math.sqrt(D)
            logits.masked_fill_(causal, float("-inf"))
            probs = torch.softmax(logits, dim=-1)
            A[:, :, r0:r1] = probs[..., :P]
            ctx[:, :, r0:r1] = probs @ v
            del logits, probs
            if want_qk:
                lb = (qb[:, :, r0:r1].float() @ kb.float().transpose(-1, -2)) / math.sqrt(D)
                lb.masked_fill_(causal, float("-inf"))
                lse[:, :, r0:r1] = torch.logsumexp(lb, dim=-1)
                del lb
        if want_qk:
            qk = (qb.contiguous(), kb[:, :, :P].contiguous(), lse, kb.contiguous(), v.to(torch.bfloat16).contiguous())
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
    return (out, qk) if qk_layer is not None else out


def cpu_synthetic_layer(S, hidden, heads, P, B, seed):
    """Reference-arm inputs, generated on the HOST (no GPU work in that arm): K,V ~ N(0,1) [B,S,HD] and a causal
    probability slice [B,H,S,P] whose rows put a decaying share of their mass on the prompt columns (like the real
    prefill's slices).  The CPU path's time depends on shapes, dtype and bit widths, not on the values."""
    rng = np.random.default_rng(seed)
    K = rng.standard_normal((B, S, hidden), dtype=np.float32)
    V = rng.standard_normal((B, S, hidden), dtype=np.float32)
    A = rng.random((B, heads, S, P), dtype=np.float32)
    i = np.arange(S, dtype=np.float32)[:, None]
    j = np.arange(P, dtype=np.float32)[None, :]
    A *= (j <= i) / np.maximum(i + 1.0, 1.0) * 1.5
    return K, V, A


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


# ----------------------------------------------------------------------------- CPU baselines
def cpu_baseline(layers_host, pcfg, bits, S, n_layers_total, B, reps=1):
    """Oracle port of the reference's CPU path (quantise every token, then select and gather, as
    the reference does) timed on the host cores over a bounded sample of layers."""
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)
    P = max(1, min(S // 5, 128))
    l0, K0, V0, A0 = layers_host[0]
    oracle.compress_layer(K0[:, :256], V0[:, :256], oracle.F32, A0[:, :, :256], oracle.F32, min(P, 51),
                          pcfg.alpha, pcfg.beta, pcfg.gamma, 1.0, pcfg.theta_h, pcfg.theta_m, bits, 0.8)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        for (l, K, V, A) in layers_host:
            oracle.compress_layer(K, V, oracle.F32, A, oracle.F32, P, pcfg.alpha, pcfg.beta, pcfg.gamma,
                                  pcfg.layer_weights[l], pcfg.theta_h, pcfg.theta_m, bits, pcfg.ratio(l))
    dt = (time.perf_counter() - t0) / reps
    per_layer = dt / len(layers_host)
    return B * S / (per_layer * n_layers_total), per_layer, oracle.num_threads()


def reference_python_leg(K, V, A, layer_idx, n_layers_total, bits, ratios=None):
    """The UNMODIFIED reference (baseline/_ref/, installed by baseline/install_ref.py) timed on host tensors:
    RealTimePrefillCompressor.compress_layer_kv_cache (src/compression/unified_compressor.py:95-172), one layer,
    every host thread torch will use.  Returns None when baseline/_ref/ is absent."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_root, "src", "compression", "unified_compressor.py")):
        return None
    try:
        sys.path.insert(0, ref_root)
        from configs.base_config import CompressionConfig as RefConfig
        from src.compression.unified_compressor import RealTimePrefillCompressor as RefCompressor
        torch.set_num_threads(os.cpu_count() or 1)
        kw = dict(num_hidden_layers=max(n_layers_total, 2), low_precision_bits=bits[0], medium_precision_bits=bits[1],
                  high_precision_bits=bits[2])
        if ratios:
            kw.update(early_layer_ratio=ratios[0], middle_layer_ratio=ratios[1], later_layer_ratio=ratios[2])
        comp = RefCompressor(RefConfig(**kw))
        Kt, Vt, At = (torch.from_numpy(np.ascontiguousarray(t)) if isinstance(t, np.ndarray) else t for t in (K, V, A))
        B, S, _ = Kt.shape
        ids = torch.zeros(B, S, dtype=torch.long)
        t0 = time.perf_counter()
        k_sel, _, _ = comp.compress_layer_kv_cache(Kt, Vt, At, ids, layer_idx)
        dt = time.perf_counter() - t0
        return {"value": B * S / (dt * n_layers_total), "unit": "tokens/s", "seconds_per_layer": dt,
                "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "kind": "reference",
                "kept_rows": int(k_sel.shape[1]),
                "sample": f"ONE layer (index {layer_idx}) of the workload, once, unmodified reference from baseline/_ref on "
                          f"CPU tensors; tokens/s = B*S / (seconds_per_layer x {n_layers_total} layers), extrapolated "
                          f"linearly over the layers"}
    except Exception as e:          # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        if ref_root in sys.path:
            sys.path.remove(ref_root)


# ----------------------------------------------------------------------------- stdout hygiene / shutdown
def _shutdown(dist):
    """Tear the process group down, but never hang the box: hard-exit if teardown stalls."""
    import threading
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        dist.destroy_process_group()
    finally:
        os._exit(0)


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


class Timer:
    """CUDA-event timing of K replays with a barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world, dist):
        self.dev, self.world, self.dist = dev, world, dist

    def run(self, fn, warmup, steps):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        ms = a.elapsed_time(b)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps


# ----------------------------------------------------------------------------- single GPU / batch-sharded measurement
def path_bytes_fn(B, S, hidden, heads, P, e=4):
    k1 = B * S * heads * P * e + B * S * 4
    k2 = B * S * 12
    return k1, k2, (lambda kept: kept * 2 * hidden * (e + e))


def measure_device_path(args, rtkv, ops, layers, cfg, bits, dev, timer, S, B, P, hidden, heads, outs, want_detail=True):
    """K1a -> K2 -> K3 over all `layers` as ONE CUDA graph per step; returns timing + roofline numbers."""
    comp = rtkv.RealTimePrefillCompressor(cfg)
    L = len(layers)

    def enqueue(events=None, side=None):
        sels = []
        main = torch.cuda.current_stream()
        for j, (l, K, V, A) in enumerate(layers):
            params, _ = comp.layer_params(l, S, ops.F32)
            mass = ops.prompt_mass(A, prompt_len=P)
            sel = ops.score_select(params, B, S, dev, mass=mass)
            ok, ov = outs[j % len(outs)]
            if side is None:
                if events is not None:
                    events[j][0].record()
                ops.quant_gather(K, V, sel, sel.labels, bits, S, zero_pad=False, out0=ok, out1=ov)
                if events is not None:
                    events[j][1].record()
            else:  # K3 of layer j overlaps K1/K2 of layer j+1 (needs its own output buffers: outs must hold L pairs)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.quant_gather(K, V, sel, sel.labels, bits, S, zero_pad=False, out0=ok, out1=ov)
            sels.append(sel)
        if side is not None:
            main.wait_stream(side)
        return sels

    for _ in range(max(1, min(args.warmup, 3))):
        sels = enqueue()
    torch.cuda.synchronize()

    def capture(events=None, two_stream=False):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev) if two_stream else None
        with torch.cuda.graph(g):
            out = enqueue(events, side)
        return g, out

    graph, sels = capture()
    ms_per_step = timer.run(graph.replay, args.warmup, args.steps)
    res = {"ms_per_step": ms_per_step, "launches_per_step": 3 * L}
    kept = [int(s.read_stats()[1]["n_selected"].sum()) for s in sels]
    k1_bytes, k2_bytes, k3_of = path_bytes_fn(B, S, hidden, heads, P)
    k3_bytes = [k3_of(n) for n in kept]
    path_bytes = sum(k3_bytes) + L * (k1_bytes + k2_bytes)
    peak, peak_src = measured_hbm_peak()
    res.update(kept=kept, retained_fraction=sum(kept) / (L * B * S), path_bytes=path_bytes, k3_bytes=sum(k3_bytes),
               whole_path={"achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           "frac_of_8TBs_nominal": path_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                           "bytes_per_step": path_bytes})
    if not want_detail:
        del graph
        return res
    # K3 launch durations inside the graph: event-record nodes around every K3 kernel
    k3_graph_ms = None
    try:
        gev = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
               for _ in layers]
        graph3, _ = capture(gev)
        for _ in range(3):
            graph3.replay()
        acc = []
        for _ in range(min(args.steps, 10)):
            graph3.replay()
            torch.cuda.synchronize()
            acc.append(sum(a.elapsed_time(b) for a, b in gev))
        k3_graph_ms = float(np.mean(acc))
        del graph3
    except Exception as e:     # pragma: no cover  (timing events in graphs not supported: keep the eager figure)
        sys.stderr.write(f"event nodes in the graph failed ({e}); K3 is timed in the eager pass only\n")
    # instrumented eager pass (same process, same data): CUDA events around every K3 launch
    n_eager = min(args.steps, 5)
    ev_k3 = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in layers]
             for _ in range(n_eager)]
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for it in range(n_eager):
        enqueue(ev_k3[it])
    t_end.record()
    torch.cuda.synchronize()
    eager_ms_per_step = t_start.elapsed_time(t_end) / n_eager
    k3_eager_ms = float(np.mean([sum(a.elapsed_time(b) for a, b in evs) for evs in ev_k3]))
    k3_ms = k3_graph_ms if k3_graph_ms is not None else k3_eager_ms
    achieved = sum(k3_bytes) / (k3_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "quant_gather_vec_kernel (K3)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                "launch_ms_avg": k3_ms / L, "bytes_per_launch_avg": sum(k3_bytes) / L,
                "k3_share_of_step": k3_ms / (ms_per_step if k3_graph_ms is not None else eager_ms_per_step),
                "measured": ("CUDA event-record nodes around every K3 kernel inside the step graph" if k3_graph_ms is not None
                             else "CUDA events around every K3 launch in an eager pass of the same run"),
                "eager_pass": {"k3_ms_per_step": k3_eager_ms, "achieved": sum(k3_bytes) / (k3_eager_ms * 1e-3) / 1e9,
                               "note": "same kernels launched from Python; includes CPU launch gaps"},
                "retained_fraction": res["retained_fraction"], "whole_path": res["whole_path"]}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("quant_gather_bytes_per_launch")
    except Exception:
        pass
    res.update(roofline=roofline, eager_ms_per_step=eager_ms_per_step)
    del graph
    return res


def measure_k1b(ops, qk, timer, steps, S, heads, P, D=128):
    """The fused Q.K_prompt^T scoring kernel (K1b: tcgen05 + TMA), the only dense contraction of the path: its HBM
    floor is streaming Q once; tensor-pipe utilisation comes from the ncu capture in profiles/."""
    q, kp, lse = qk[:3]
    B = q.shape[0]
    scale = 1.0 / math.sqrt(D)
    fn = lambda: ops.prompt_mass_qk_partials(q, kp, lse, scale, causal=True)
    ms = timer.run(fn, 3, max(steps, 10))
    q_bytes = q.numel() * 2 + lse.numel() * 4 + B * heads * S * 4
    flops = 2.0 * B * heads * S * P * D
    peak, tf_peak, src = measured_peaks()
    return {"kernel": "prompt_mass_qk_kernel (K1b, tcgen05.mma + TMA, bf16)", "ms_per_launch": ms,
            "bound": "hbm", "achieved": q_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": q_bytes / (ms * 1e-3) / 1e9 / peak, "bytes_per_launch": q_bytes,
            "tflops": flops / (ms * 1e-3) / 1e12, "tflops_frac_of_bf16_peak": flops / (ms * 1e-3) / 1e12 / tf_peak,
            "note": "one layer's Q [B,H,S,128] bf16 + LSE streamed once, P=128 prompt keys resident; replaces K1a's read of "
                    "the materialised probability slice (2x the bytes) when the caller has Q/K/LSE instead of S x S probabilities"}


def measure_flash(ops, qk, timer, steps, S, heads, P, D=128):
    """rtkv_flash_prefill: the hand-written tcgen05 + TMA prefill attention that produces the path's score input (attention
    output + LSE + per-head prompt mass in one launch).  Tensor-core bound: causal FLOPs / time against the
    measured cuBLAS bf16 peak."""
    q, _, lse_ref, k, v = qk
    B = q.shape[0]
    scale = 1.0 / math.sqrt(D)
    fn = lambda: ops.flash_prefill(q, k, v, scale, causal=True, prompt_len=P)
    ms = timer.run(fn, 2, max(3, min(steps, 5)))
    out, lse, hs = fn()
    flops = 4.0 * B * heads * S * S * D / 2.0
    _, tf_peak, src = measured_peaks()
    sustained = measured_bf16_sustained()
    achieved = flops / (ms * 1e-3) / 1e12
    return {"kernel": "rtkv_flash_prefill (tcgen05.mma + TMA, bf16, causal, LSE + fused prompt mass; kernel chosen by key count)", "ms_per_launch": ms,
            "bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
            "frac": achieved / tf_peak, "peak_source": src + ", burst figure: the kernel is timed alone",
            "peak_sustained": sustained, "frac_of_sustained": (achieved / sustained) if sustained else None,
            "lse_max_abs_diff_vs_torch_fp32": float((lse - lse_ref).abs().max()),
            "note": "one layer's causal attention [B,H,S,128] x [B,H,S,128]; FLOPs = 4*B*H*S^2*D/2; replaces the library "
                    "attention + separate scoring kernel of the model hook"}


# ----------------------------------------------------------------------------- head-sharded mode
def bench_head_sharded(args, rtkv, ops, layers, cfg, config, dev, rank, world, L, S, B, P, hidden, heads, bits, timer, dist):
    """BASELINE configs[2]: one long sequence, KV heads split over the ranks (strong scaling).
    mode fused:       shard_stats (A,K,V streamed once, records pushed to every rank) -> shard_combine -> K2 -> K3
    mode collective:  K1a partials -> all-gather -> K2 -> row min/max -> MIN all-reduce -> K3
    mode p2p:         K1a partials -> all-gather -> K2 -> K3 with in-kernel min/max mailboxes
    No host synchronisation in any mode."""
    from rtkv import dist as rd
    comp = rtkv.RealTimePrefillCompressor(cfg)
    mode = args.shard_mode
    if os.environ.get("RTKV_SHARD_VARIANT"):      # bisecting aid (timing experiments only; 2/3 give wrong results)
        ops.lib.rtkv_debug_shard_variant(int(os.environ["RTKV_SHARD_VARIANT"]))
    sharded = rd.HeadShardedCompressor(comp, heads, mode=mode)
    try:
        if mode == "fused":
            sharded._exchange(B, S, dev)
        elif mode == "p2p":
            sharded._mailbox(B, S, dev)
    except Exception as e:            # symmetric memory unavailable: fall back to the collective scheme
        sys.stderr.write(f"symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL collective scheme\n")
        mode = "collective"
        sharded = rd.HeadShardedCompressor(comp, heads, mode=mode)
    w = hidden // world
    outs = [(torch.empty(B, S, w, device=dev), torch.empty(B, S, w, device=dev)) for _ in range(2)]

    def step():
        sels = []
        for j, (l, K, V, A) in enumerate(layers):
            sels.append(sharded.compress_layer_async(K, V, A, l, outs[j & 1][0], outs[j & 1][1]))
        return sels

    for _ in range(max(1, min(args.warmup, 3))):
        sels = step()
    torch.cuda.synchronize()
    dist.barrier()
    launch_mode = "eager launches"
    graph = None
    try:                                   # NCCL collectives can be captured too; fall back to eager if not
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        graph.replay()
        torch.cuda.synchronize()
        launch_mode = "one CUDA graph per step"
    except Exception as e:                 # pragma: no cover
        sys.stderr.write(f"graph capture of the head-sharded step failed ({e}); timing eager launches\n")
        graph = None
        torch.cuda.synchronize()
    run = graph.replay if graph is not None else step
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    ms_per_step = timer.run(run, args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    xerr = bool(sharded.exchange.error()) if mode == "fused" else False
    kept = [int(s.read_stats()[1]["n_selected"].sum()) for s in sels]
    e = 4
    # algorithmic bytes: every rank's share of A once + kept K/V rows once in, once out.  The exact
    # whole-row scale forces a second read of the local K/V rows (min/max before the exchange); it is NOT counted.
    path_bytes = sum(n * 2 * hidden * 2 * e for n in kept) + L * (B * S * heads * P * e + B * S * 16)
    peak, peak_src = measured_hbm_peak()
    launches = {"fused": 4, "collective": 5, "p2p": 4}[mode]
    line = {"metric": METRIC, "value": B * S / (ms_per_step * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "clocks": clocks, "gpu_launches": launches * L * args.steps, "launch_mode": launch_mode,
            "roofline": {"bound": "hbm", "kernel": "whole path, all ranks", "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                         "peak_source": peak_src + f" x {world} GPUs", "traffic": None,
                         "retained_fraction": sum(kept) / (L * B * S),
                         "note": "algorithmic bytes of the UNSHARDED path / step time; the sharded path additionally re-reads "
                                 "the local K/V rows once (min/max before the exchange), which is not counted as useful"},
            "exchange": {"fused": "ONE per layer, in-kernel: every token's (kmin,kmax,vmin,vmax, mass partials) record is "
                                  "peer-stored into every rank's symmetric-memory buffer by the statistics kernel; epoch flags; "
                                  "no NCCL call on the data path",
                         "collective": "two NCCL calls per layer: all_gather_into_tensor(mass partials) + MIN all_reduce(row min/max)",
                         "p2p": "NCCL all-gather of the mass partials + row min/max mailboxes inside the quantise kernel"}[mode],
            "exchange_error": xerr,
            "exchange_bytes_per_layer_per_rank": {"sent": B * S * 32 * (world - 1) if mode == "fused" else None}}
    if not args.no_e2e:
        line["e2e"] = e2e_head_sharded(args, sharded, layers, dev, world, L, S, B, w, dist)
    return line, run, graph


def e2e_head_sharded(args, sharded, layers, dev, world, L, S, B, w, dist):
    """End to end for the head-sharded job: every rank's K, V and attention slices start in pinned HOST memory and its slice of
    the compressed K/V ends there.  Per layer: H2D of the three input slices (copy stream, one layer ahead), the sharded layer
    through the public call HeadShardedCompressor.compress_layer_kv_cache (it synchronises once: S' is a shape), D2H of the
    [B,S',w] results (download stream).  All N ranks run it together (the exchange is per layer); max over ranks."""
    n_host = max(1, min(L, 8))
    host_in = []
    for j in range(n_host):
        l, K, V, A = layers[j * L // n_host]
        hk, hv, ha = (torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (K, V, A))
        hk.copy_(K, non_blocking=True); hv.copy_(V, non_blocking=True); ha.copy_(A, non_blocking=True)
        host_in.append((hk, hv, ha))
    torch.cuda.synchronize()
    _, K0, V0, A0 = layers[0]
    dbuf = [(torch.empty_like(K0), torch.empty_like(V0), torch.empty_like(A0)) for _ in range(2)]
    hout = [(torch.empty(B, S, w, dtype=K0.dtype, pin_memory=True), torch.empty(B, S, w, dtype=K0.dtype, pin_memory=True))
            for _ in range(2)]
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]          # device input buffers consumed by the layer
    ev_down = [torch.cuda.Event() for _ in range(2)]

    def upload(j):
        hk, hv, ha = host_in[j * n_host // L]
        dk, dv, da = dbuf[j & 1]
        with torch.cuda.stream(up):
            up.wait_event(ev_free[j & 1])
            da.copy_(ha, non_blocking=True); dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True)
            ev_up[j & 1].record(up)

    def step():
        d2h = 0
        for e in ev_free + ev_down:
            e.record(main)
        upload(0)
        for j, (l, _, _, _) in enumerate(layers):
            if j + 1 < L:
                upload(j + 1)
            main.wait_event(ev_up[j & 1])
            dk, dv, da = dbuf[j & 1]
            k_sel, v_sel, sel = sharded.compress_layer_kv_cache(dk, dv, da, l)
            ev_free[j & 1].record(main)
            ok, ov = hout[j & 1]
            with torch.cuda.stream(down):
                down.wait_stream(main)
                down.wait_event(ev_down[j & 1])
                ms = k_sel.shape[1]
                ok[:, :ms].copy_(k_sel, non_blocking=True); ov[:, :ms].copy_(v_sel, non_blocking=True)
                ev_down[j & 1].record(down)
            k_sel.record_stream(down); v_sel.record_stream(down)
            d2h += 2 * k_sel.numel() * k_sel.element_size()
        torch.cuda.synchronize()
        return d2h

    step()
    dist.barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d2h = step()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = L * sum(x.numel() * x.element_size() for x in host_in[0])
    return {"value": B * S / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
            "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "per_rank_h2d_bytes_per_step": h2d,
            "host_input_sets": n_host,
            "api": "HeadShardedCompressor.compress_layer_kv_cache per rank on slices that start in pinned host memory and end there "
                   "(H2D one layer ahead on a copy stream, D2H on a download stream); bytes are the totals over all ranks"}


# ----------------------------------------------------------------------------- main
def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="rtkv", choices=["rtkv", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="default c3 = the metric's configuration (7B, 32 layers x 32768 tokens)")
    ap.add_argument("--layers", type=int, default=0, help="override the number of layers (setup time)")
    ap.add_argument("--bits", default="", help="low,mid,high bit widths, e.g. 8,8,8 (configs[3] precision sweep)")
    ap.add_argument("--ablation", default="", choices=["", "quant_only"],
                    help="quant_only: propagation ratios 1.0, every token kept (configs[4] ablation)")
    ap.add_argument("--sharding", default="auto", choices=["auto", "heads", "batch"],
                    help="N>1: auto = KV heads for c3 (strong scaling), batch otherwise (weak scaling)")
    ap.add_argument("--shard-mode", default="fused", choices=["fused", "collective", "p2p"],
                    help="head-sharded exchange scheme (rtkv/dist.py)")
    ap.add_argument("--p2p", action="store_true", help="alias of --shard-mode p2p")
    ap.add_argument("--no-pdl", action="store_true", help="launch without programmatic dependent launch (A/B)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary measurements (16/8/4, K1b, batch-sharded)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "rtkv" else args.warmup
    if args.p2p:
        args.shard_mode = "p2p"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, L, hidden, heads, inter, S, B, bits = WORKLOADS[args.workload]
    if args.layers > 0:
        L = args.layers
    if args.bits:
        bits = tuple(int(x) for x in args.bits.split(","))
    n_gpus = args.gpus if args.impl == "reference" else world
    head_sharded = (args.sharding == "heads" or (args.sharding == "auto" and args.workload == "c3")) and n_gpus > 1
    P = max(1, min(S // 5, 128))
    sample_layers = sorted({0, L // 3, (2 * L) // 3, L - 1})
    ratios = (1.0, 1.0, 1.0) if args.ablation == "quant_only" else None
    pcfg = PathConfig(max(L, 2), bits, ratios)
    config = {"workload": desc, "layers": L, "seq_len": S, "batch_per_gpu": B, "kv_width": hidden, "heads": heads,
              "prompt_len": P, "bits_low_mid_high": list(bits), "theta_h": pcfg.theta_h, "theta_m": pcfg.theta_m,
              "ratios": [pcfg.early_layer_ratio, pcfg.middle_layer_ratio, pcfg.later_layer_ratio],
              "kv_dtype": "float32", "attention_input": f"[B,H,S,P={P}] probability slice, float32",
              "sharding": ("KV heads over the GPUs, same problem as N=1 (strong scaling); one in-kernel exchange per layer"
                           if head_sharded else
                           "batch (one sequence set per GPU, no collective)" if n_gpus > 1 else "single GPU"),
              "l2": "per-step working set (inputs+outputs, all layers) far exceeds the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU only: no rtkv, no CUDA)
    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun pins OMP_NUM_THREADS=1; the reference arm uses every host core (rank 0 runs alone)
        import oracle
        oracle.set_num_threads(os.cpu_count() or 1)
        K, V, A = cpu_synthetic_layer(S, hidden, heads, P, B, 1234)
        host = [(l, K, V, A) for l in sample_layers]
        vals = []
        for it in range(args.warmup + args.steps):
            v, per_layer, cores = cpu_baseline(host, pcfg, bits, S, L, B)
            if it >= args.warmup:
                vals.append((v, per_layer))
        value = float(np.mean([v for v, _ in vals]))
        per_layer = float(np.mean([p for _, p in vals]))
        refpy = reference_python_leg(K, V, A, sample_layers[len(sample_layers) // 2], L, bits, ratios)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_layer * len(sample_layers) * 1e3, "higher_is_better": True,
                "scaling": "strong" if head_sharded else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
                                 "sample": f"layers {sample_layers} of {L} per step (one synthetic layer's tensors, generated on "
                                           f"the host, with each sampled layer's weight and depth-band ratio), C/OpenMP port of "
                                           f"the reference's CPU path (oracle/), per-layer time x {L} layers; no GPU work and no "
                                           f"rtkv import in this arm",
                                 "reference_python": refpy},
                "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    # ------------------------------------------------------------------ rtkv arm (GPU)
    import rtkv
    from rtkv import ops
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    timer = Timer(dev, world, dist)
    ckw = {} if ratios is None else {"early_layer_ratio": 1.0, "middle_layer_ratio": 1.0, "later_layer_ratio": 1.0}
    cfg = rtkv.CompressionConfig(num_hidden_layers=max(L, 2), num_attention_heads=heads, hidden_size=hidden,
                                 high_precision_bits=bits[2], medium_precision_bits=bits[1],
                                 low_precision_bits=bits[0], **ckw)
    if args.no_pdl:
        from rtkv import _lib as _rl
        _rl.lib.rtkv_debug_pdl(0)
    secondary = not args.no_secondary
    want_qk = secondary and world == 1 and hidden // heads == 128 and L > 1
    gen = llama_prefill_inputs(L, hidden, heads, inter, S, B, P, 1234 + (0 if head_sharded else rank), dev,
                               head_slice=(rank, world) if head_sharded else None,
                               qk_layer=(L // 2 if want_qk else None))
    layers, qk = gen if want_qk else (gen, None)
    torch.cuda.synchronize()

    if head_sharded:
        line, run, graph = bench_head_sharded(args, rtkv, ops, layers, cfg, config, dev, rank, world, L, S, B, P, hidden,
                                              heads, bits, timer, dist)
        # a live CUDA graph must be dropped before the process group goes away
        del run, graph
        torch.cuda.synchronize()
        if secondary:      # secondary key: batch sharding (weak scaling, no exchange), one 4k sequence per GPU
            del layers
            torch.cuda.empty_cache()
            d2, L2, h2, hd2, i2, S2, B2, bits2 = WORKLOADS["c2"]
            P2 = max(1, min(S2 // 5, 128))
            lay2 = llama_prefill_inputs(L2, h2, hd2, i2, S2, B2, P2, 4321 + rank, dev)
            cfg2 = rtkv.CompressionConfig(num_hidden_layers=L2, num_attention_heads=hd2, hidden_size=h2,
                                          high_precision_bits=bits2[2], medium_precision_bits=bits2[1], low_precision_bits=bits2[0])
            outs2 = [(torch.empty_like(K), torch.empty_like(V)) for (_, K, V, _) in lay2]
            r2 = measure_device_path(args, rtkv, ops, lay2, cfg2, bits2, dev, timer, S2, B2, P2, h2, hd2, outs2, want_detail=False)
            line["batch_sharded"] = {"workload": d2 + ", one sequence per GPU, no exchange", "scaling": "weak",
                                     "ms_per_step": r2["ms_per_step"], "value": world * B2 * S2 / (r2["ms_per_step"] * 1e-3),
                                     "unit": "tokens/s", "whole_path_frac_per_gpu": r2["whole_path"]["frac"]}
        if rank == 0:
            _emit(line)
        _shutdown(dist)
        return

    # ---- single GPU (or batch-sharded: every rank its own sequences, no collective) ----
    n_out = len(layers) if S * B <= 8192 * 4 else 2      # long sequences: two rotating output pairs (K3 writes are stream-ordered)
    outs = [(torch.empty_like(layers[0][1]), torch.empty_like(layers[0][2])) for _ in range(n_out)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = measure_device_path(args, rtkv, ops, layers, cfg, bits, dev, timer, S, B, P, hidden, heads, outs)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = res["ms_per_step"]
    value = world * B * S / (ms_per_step * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": 3 * L * args.steps, "roofline": res["roofline"],
            "launch_mode": "one CUDA graph per step (3 kernels x %d layers, single stream%s)" % (
                L, "" if args.no_pdl else ", programmatic dependent launch"),
            "eager_ms_per_step": res["eager_ms_per_step"],
            "reference_recorded": {"tokens_per_s": 173, "note": "BASELINE.md: derived from the reference's recorded "
                                   "runs (23.7 s per 4096-token prefill, bits 16/8/4, unnamed GPU); not a published number"}}

    if secondary and world == 1:
        # the reference's DEFAULT widths 16/8/4 on the same inputs: the budget binds, K2 walks, K3 gathers
        try:
            b2 = (4, 8, 16)
            cfg2 = rtkv.CompressionConfig(num_hidden_layers=max(L, 2), num_attention_heads=heads, hidden_size=hidden,
                                          high_precision_bits=16, medium_precision_bits=8, low_precision_bits=4, **ckw)
            r2 = measure_device_path(args, rtkv, ops, layers, cfg2, b2, dev, timer, S, B, P, hidden, heads, outs)
            line["bits_16_8_4"] = {"note": "same inputs with the reference's default widths (configs/base_config.py:33-35): the "
                                           "budget binds, so the selection walk and the gather are exercised",
                                   "ms_per_step": r2["ms_per_step"], "value": B * S / (r2["ms_per_step"] * 1e-3),
                                   "retained_fraction": r2["retained_fraction"], "whole_path": r2["whole_path"],
                                   "k3": {k: r2["roofline"][k] for k in ("achieved", "frac", "launch_ms_avg", "bytes_per_launch_avg",
                                                                         "k3_share_of_step")}}
        except Exception as e:      # pragma: no cover
            line["bits_16_8_4"] = {"error": f"{type(e).__name__}: {e}"}
        if qk is not None:
            try:
                line["k1b_scoring"] = measure_k1b(ops, qk, timer, args.steps, S, heads, P)
            except Exception as e:  # pragma: no cover
                line["k1b_scoring"] = {"error": f"{type(e).__name__}: {e}"}
            try:
                line["flash_prefill"] = measure_flash(ops, qk, timer, args.steps, S, heads, P)
            except Exception as e:  # pragma: no cover
                line["flash_prefill"] = {"error": f"{type(e).__name__}: {e}"}
    del qk

    # ------------------------------------------------------------------ e2e: public API, host buffers
    if not args.no_e2e:
        import psutil
        comp = rtkv.RealTimePrefillCompressor(cfg)
        per_layer_bytes = sum(t.numel() * 4 for t in layers[0][1:])
        avail = psutil.virtual_memory().available
        n_host = L if per_layer_bytes * L * 1.3 + (4 << 30) < avail else max(1, int((avail * 0.5) // per_layer_bytes))
        if per_layer_bytes * L > (16 << 30):     # pinning host memory runs at ~1 GB/s on the box: bound the setup time
            n_host = min(n_host, max(1, (12 << 30) // per_layer_bytes))
        n_host = max(1, min(n_host, L))
        host_in = []
        for j in range(n_host):          # pinned host copies of the inputs (all layers when host memory allows)
            l, K, V, A = layers[j * L // n_host]
            hk, hv, ha = (torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (K, V, A))
            hk.copy_(K, non_blocking=True); hv.copy_(V, non_blocking=True); ha.copy_(A, non_blocking=True)
            host_in.append((hk, hv, ha))
        torch.cuda.synchronize()
        ids = torch.zeros(B, S, dtype=torch.long)
        ko = [torch.empty(B * S * hidden, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        vo = [torch.empty(B * S * hidden, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        e2e_steps = max(1, min(args.steps, 3 if S * B > 16384 else 5))

        def e2e_step():
            d2h = 0
            for j, (l, _, _, _) in enumerate(layers):
                K, V, A = host_in[j * n_host // L]
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
        h2d = L * per_layer_bytes
        line["e2e"] = {"value": world * B * S / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                       "steps": e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "host_input_sets": n_host,
                       "api": "RealTimePrefillCompressor.compress_layer_kv_cache on pinned host tensors "
                              "(rtkv_compress_layer_host), synchronous per layer; every layer call copies its K, V and "
                              "attention slice up and its compacted result down inside the timed region" +
                              ("" if n_host == L else f"; host memory holds {n_host} distinct layers' inputs, reused in "
                                                      f"depth order for the {L} layer calls")}
        del host_in, ko, vo

    # ------------------------------------------------------------------ CPU baselines (rank 0, N=1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host = [(l, K.cpu().numpy(), V.cpu().numpy(), A.cpu().numpy()) for (l, K, V, A) in layers if l in sample_layers]
        v, per_layer, cores = cpu_baseline(host, pcfg, bits, S, L, B)
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                                "sample": f"layers {sample_layers} of {L}, once each, C/OpenMP port of the reference's "
                                          f"CPU path (oracle/): {per_layer * 1e3:.1f} ms per layer x {L} layers"}
        mid = host[len(host) // 2]
        line["cpu_baseline"]["reference_python"] = reference_python_leg(mid[1], mid[2], mid[3], mid[0], L, bits, ratios)
    if rank == 0:
        _emit(line)
    if world > 1:
        _shutdown(dist)


if __name__ == "__main__":
    main()
