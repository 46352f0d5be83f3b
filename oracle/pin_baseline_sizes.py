#!/usr/bin/env python
"""Pin parity against the REAL reference at BASELINE sizes.

Runs ONLY in the build container (imports the unmodified reference from /root/reference, read-only).
For BASELINE configs[0] ([1,32,2048,128] -> K/V [1,2048,4096], layers 0/15/31, bits 8/4/2 and 16/8/4), one
[1,4096,4096] layer (configs[1] shape) and one [1,32768,4096] layer (the metric's 32k shape) it
  1. builds the inputs with `baseline_inputs()` below — integer draws turned into floats with exact
     elementwise arithmetic only, so the GPU box regenerates them bit for bit (checked by CRC),
  2. runs `RealTimePrefillCompressor.compress_layer_kv_cache` of the reference on CPU
     (src/compression/unified_compressor.py:95-172; argsort made stable = the lowest-index tie contract),
  3. stores a compact digest in tests/golden/baseline_<name>.npz: scores, labels, mask, per-row CRC32 of
     k_sel / v_sel and 64 sampled rows of each,
  4. compares the CPU oracle with the reference on the same inputs and records, per case, the number of
     selection flips, label flips, score ulp distances and the decision margins of the selection
     (tests/golden/PINNING_BASELINE.json).

Decision margin: the selection keeps, per precision class, a prefix of the tokens in descending score order
(selective_propagation.py:113-131).  A 1-ulp difference in a score (the H*P-term prompt-mass sum is the one stage
whose operation order ATen does not define) can only flip the selection if it reorders a kept token and a dropped
token of the same class, i.e. if the class boundary gap — lowest kept score minus highest dropped score — is within
the score error.  The digest stores that gap in ulps for every class that has a boundary.

Usage:  python oracle/pin_baseline_sizes.py
"""
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, S, layer, bits(low,mid,high), seed
    ("c1_l0_842", 2048, 0, (2, 4, 8), 101),
    ("c1_l15_842", 2048, 15, (2, 4, 8), 102),
    ("c1_l31_842", 2048, 31, (2, 4, 8), 103),
    ("c1_l0_1684", 2048, 0, (4, 8, 16), 104),
    ("c1_l15_1684", 2048, 15, (4, 8, 16), 105),
    ("c1_l31_1684", 2048, 31, (4, 8, 16), 106),
    ("c2_l25_842", 4096, 25, (2, 4, 8), 107),
    ("c2_l12_1684", 4096, 12, (4, 8, 16), 108),
    ("c3_l28_842", 32768, 28, (2, 4, 8), 109),
    ("c3_l20_1684", 32768, 20, (4, 8, 16), 110),
    # BASELINE configs[3] (C4): Llama-2-13B shapes, 16384 tokens, the four bit settings of the sweep; (L, H, HD) as 6th field
    ("c4_l35_888", 16384, 35, (8, 8, 8), 111, (40, 40, 5120)),
    ("c4_l30_444", 16384, 30, (4, 4, 4), 112, (40, 40, 5120)),
    ("c4_l20_222", 16384, 20, (2, 2, 2), 113, (40, 40, 5120)),
    ("c4_l14_842", 16384, 14, (2, 4, 8), 114, (40, 40, 5120)),
    # BASELINE configs[4] (C5): one rank's share of the 64-sequence batch (8 sequences of 4096 tokens, 7B shapes); 16/8/4 on a
    # later layer so that the rows keep different numbers of tokens and the zero padding to max_selected is exercised; 7th field = B
    ("c5_b8_l25_1684", 4096, 25, (4, 8, 16), 115, (32, 32, 4096), 8),
    # configs[0] shape in the deployment dtypes: K, V and the attention slice rounded to bf16 / f16 (8th field), every reference
    # operation then runs in that dtype (dynamic_quantization.py:62-126 on half tensors)
    ("c1_l15_842_bf16", 2048, 15, (2, 4, 8), 116, (32, 32, 4096), 1, "bf16"),
    ("c1_l31_888_f16", 2048, 31, (8, 8, 8), 117, (32, 32, 4096), 1, "f16"),         # the budget binds: selection + gather in f16
    ("c3_l28_888_bf16", 32768, 28, (8, 8, 8), 118, (32, 32, 4096), 1, "bf16"),      # the metric's 32k shape in the deployment dtype
]
H, HD, L = 32, 4096, 32


def batch_of(case):
    return case[6] if len(case) > 6 else 1


def batched_inputs(S, seed, heads, width, B):
    """B independent sequences (seed, seed + 1000, ...) stacked along the batch dimension."""
    if B == 1:
        return baseline_inputs(S, seed, heads=heads, width=width)
    parts = [baseline_inputs(S, seed + 1000 * b, heads=heads, width=width) for b in range(B)]
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]), parts[0][3])


def dtype_of(case):
    return case[7] if len(case) > 7 else "f32"


def as_dtype(a, name):
    """float32 numpy array -> (array in the storage the oracle takes: float32, or uint16 bit patterns; oracle dtype code).
    Rounding to the half types is torch's (round to nearest even): deterministic, checked by the CRC in the fixture."""
    import oracle
    if name == "f32":
        return a, oracle.F32
    t = torch.from_numpy(a).to(torch.bfloat16 if name == "bf16" else torch.float16)
    return oracle.from_torch(t)


def model_of(case):
    """(L, H, HD) of a CASES entry: the 7B shapes unless the entry names its own."""
    return case[5] if len(case) > 5 else (L, H, HD)


def baseline_inputs(S, seed, heads=H, width=HD):
    """K, V [1,S,width] and the prompt columns of a causal attention-like matrix A [1,heads,S,P], float32.
    Platform-exact: PCG64 integer draws; every float is produced by exact or single correctly-rounded
    elementwise IEEE operations (no reductions, no transcendental functions)."""
    rng = np.random.default_rng(seed)
    P = max(1, min(S // 5, 128))

    def bell(shape, shift):
        acc = np.zeros(shape, dtype=np.int32)
        for _ in range(4):                                   # sum of 4 uniform 20-bit integers: exact in int32
            acc += rng.integers(0, 1 << 20, size=shape, dtype=np.int32)
        return ((acc - (1 << 21)).astype(np.float32) * np.float32(2.0 ** -19)) + np.float32(shift)   # ~N(shift, 0.58^2), exact

    K = bell((1, S, width), 0.0)
    V = bell((1, S, width), 0.25)
    u = rng.integers(0, 1 << 24, size=(1, heads, S, P), dtype=np.int32).astype(np.float32) * np.float32(2.0 ** -24)
    i = np.arange(S, dtype=np.float32)[:, None]
    j = np.arange(P, dtype=np.float32)[None, :]
    decay = np.float32(1.5) / (i + np.float32(1.0))          # one IEEE division per row
    A = u * (decay * (j <= i).astype(np.float32))[None, None]
    # a few "sink" rows that put a large share on the prompt (spreads the min-max normalised mass over [0,1])
    hot = rng.integers(0, S, size=max(4, S // 64))
    A[:, :, hot, :] *= np.float32(8.0)
    return K, V, A, P


def crc_rows(x):
    x = np.ascontiguousarray(x)
    return np.array([zlib.crc32(x[b, r].tobytes()) for b in range(x.shape[0]) for r in range(x.shape[1])],
                    dtype=np.uint32).reshape(x.shape[0], x.shape[1])


def crc_all(x):
    return zlib.crc32(np.ascontiguousarray(x).tobytes())


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-(2 ** 31)) - ai, ai)
    bi = np.where(bi < 0, np.int64(-(2 ** 31)) - bi, bi)
    return np.abs(ai - bi)


def boundary_margins(scores, labels, mask):
    """Per class with both kept and dropped tokens: (lowest kept score) - (highest dropped score) in ulps of the
    larger magnitude; negative would mean the class is not a prefix (never happens for monotone labels)."""
    out = {}
    for c in range(3):
        kept = scores[(labels == c) & (mask == 1)]
        drop = scores[(labels == c) & (mask == 0)]
        if kept.size and drop.size:
            lo, hi = np.float32(kept.min()), np.float32(drop.max())
            out[str(c)] = int(ulp_diff(np.array([lo]), np.array([hi]))[0]) * (1 if lo >= hi else -1)
    return out


def survey_large(n, oracle, CompressionConfig, RealTimePrefillCompressor, stable_argsort):
    """Oracle vs reference selection flips where the budget binds at the LONG shapes (16384 and 32768 tokens, 7B widths):
    tally only (tests/golden/PINNING_BASELINE.json: flip_survey_large), about 20 s of reference time per layer."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle.set_num_threads(os.cpu_count() or 1)
    sv = dict(cases=0, selection_flips=0, label_flips=0, outputs_identical=0, score_max_ulp=0, tokens=0, kept=0,
              min_class_boundary_margin_ulps=None, margin_histogram_ulps={"<=2": 0, "3-16": 0, "17-256": 0, ">256": 0}, shapes=[])
    srng = np.random.default_rng(4242)
    for t in range(n):
        S = [16384, 32768][t % 2]
        bits = [(8, 8, 8), (4, 8, 16), (4, 4, 4), (8, 8, 8)][t % 4]
        layer = int(srng.integers(9, 32))
        K, V, A, P = baseline_inputs(S, 9000 + t)
        cfg = CompressionConfig(num_hidden_layers=L, num_attention_heads=H, hidden_size=HD,
                                high_precision_bits=bits[2], medium_precision_bits=bits[1], low_precision_bits=bits[0])
        comp = RealTimePrefillCompressor(cfg)
        with stable_argsort():
            k_sel, v_sel, info = comp.compress_layer_kv_cache(torch.from_numpy(K), torch.from_numpy(V), torch.from_numpy(A),
                                                              torch.zeros(1, S, dtype=torch.long), layer)
        scores = comp.importance_tracker.layer_scores[layer].numpy()
        labels = info["quantization_info"]["bit_assignments"].astype(np.uint8)
        mask = info["propagation_info"]["selection_mask"].numpy().astype(np.uint8)
        o = oracle.compress_layer(K, V, oracle.F32, A, oracle.F32, P, cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer],
                                  cfg.theta_h, cfg.theta_m, bits, info["propagation_info"]["propagation_ratio"])
        flips = int((mask != o["mask"]).sum())
        sv["cases"] += 1
        sv["selection_flips"] += flips
        sv["label_flips"] += int((labels != o["labels"]).sum())
        sv["outputs_identical"] += int(flips == 0 and np.array_equal(k_sel.numpy(), o["k_sel"]) and np.array_equal(v_sel.numpy(), o["v_sel"]))
        sv["score_max_ulp"] = max(sv["score_max_ulp"], int(ulp_diff(scores, o["scores"]).max()))
        sv["tokens"] += S
        sv["kept"] += int(mask.sum())
        sv["shapes"].append([S, list(bits), layer, int(mask.sum())])
        for c, m in boundary_margins(scores[0], labels[0], mask[0]).items():
            cur = sv["min_class_boundary_margin_ulps"]
            sv["min_class_boundary_margin_ulps"] = m if cur is None else min(cur, m)
            key = "<=2" if m <= 2 else "3-16" if m <= 16 else "17-256" if m <= 256 else ">256"
            sv["margin_histogram_ulps"][key] += 1
        print("survey-large", t, S, bits, layer, "kept", int(mask.sum()), "flips", flips, flush=True)
    path = os.path.join(GOLD, "PINNING_BASELINE.json")
    old = json.load(open(path))
    old["flip_survey_large"] = sv
    with open(path, "w") as f:
        json.dump(old, f, indent=1)
    print(json.dumps(sv, indent=1))


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    import oracle
    from configs.base_config import CompressionConfig
    from src.compression.unified_compressor import RealTimePrefillCompressor
    from oracle.pin_against_reference import stable_argsort

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle.set_num_threads(os.cpu_count() or 1)
    tally = {"torch": torch.__version__, "cases": {}}
    rng_rows = np.random.default_rng(7)
    if "--survey-large" in sys.argv:                          # flip tally at 16k / 32k tokens only, merged into the JSON
        survey_large(int(sys.argv[sys.argv.index("--survey-large") + 1]), oracle, CompressionConfig, RealTimePrefillCompressor, stable_argsort)
        return
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None     # e.g. --only c4: these cases, no survey
    for case in CASES:
        name, S, layer, bits, seed = case[:5]
        Lc, Hc, HDc = model_of(case)
        if only and not name.startswith(only):
            continue
        t0 = time.time()
        Bc = batch_of(case)
        dname = dtype_of(case)
        K, V, A, P = batched_inputs(S, seed, Hc, HDc, Bc)
        (K, dt), (V, _), (A, adt) = as_dtype(K, dname), as_dtype(V, dname), as_dtype(A, dname)
        cfg = CompressionConfig(num_hidden_layers=Lc, num_attention_heads=Hc, hidden_size=HDc,
                                high_precision_bits=bits[2], medium_precision_bits=bits[1], low_precision_bits=bits[0])
        comp = RealTimePrefillCompressor(cfg)
        ids = torch.zeros(Bc, S, dtype=torch.long)
        with stable_argsort():
            k_sel, v_sel, info = comp.compress_layer_kv_cache(oracle.to_torch(K, dt), oracle.to_torch(V, dt), oracle.to_torch(A, adt),
                                                              ids, layer)
        t_ref = time.time() - t0
        scores = comp.importance_tracker.layer_scores[layer].numpy()
        labels = info["quantization_info"]["bit_assignments"].astype(np.uint8)
        mask = info["propagation_info"]["selection_mask"].numpy().astype(np.uint8)
        k_sel, v_sel = oracle.from_torch(k_sel)[0], oracle.from_torch(v_sel)[0]
        ratio = info["propagation_info"]["propagation_ratio"]
        # ---- oracle on the same inputs
        o = oracle.compress_layer(K, V, dt, A, adt, P, cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer],
                                  cfg.theta_h, cfg.theta_m, bits, ratio)
        ud = ulp_diff(scores, o["scores"])
        sel_flips = int((mask != o["mask"]).sum())
        lab_flips = int((labels != o["labels"]).sum())
        same_out = bool(sel_flips == 0 and np.array_equal(k_sel, o["k_sel"]) and np.array_equal(v_sel, o["v_sel"]))
        margins = boundary_margins(scores[0], labels[0], mask[0])
        for b in range(1, Bc):                                  # smallest margin over the batch rows
            for c, mg in boundary_margins(scores[b], labels[b], mask[b]).items():
                margins[c] = min(margins.get(c, mg), mg)
        # distance of every score to the nearest label threshold, in ulps (a label flip needs it within the score error)
        th = np.array([np.float32(cfg.theta_h), np.float32(cfg.theta_m)])
        thr_ulps = int(min(ulp_diff(scores.ravel(), np.full(scores.size, t, dtype=np.float32)).min() for t in th))
        kept = int(mask.sum())
        sample = np.sort(rng_rows.choice(k_sel.shape[1], size=min(64, k_sel.shape[1]), replace=False))
        meta = dict(name=name, S=S, layer=layer, bits=list(bits), seed=seed, H=Hc, HD=HDc, L=Lc, B=Bc, dtype=dname, P=P, ratio=ratio,
                    layer_weight=cfg.layer_weights[layer], alpha=cfg.alpha, beta=cfg.beta, gamma=cfg.gamma,
                    theta_h=cfg.theta_h, theta_m=cfg.theta_m, kept=kept, compressed_shape=list(k_sel.shape),
                    input_crc=dict(K=crc_all(K), V=crc_all(V), A=crc_all(A)),
                    precision_stats=info["precision_stats"], importance_stats=info["importance_stats"],
                    reference_seconds=t_ref)
        np.savez_compressed(os.path.join(GOLD, f"baseline_{name}.npz"),
                            scores=scores, labels=labels, mask=mask, k_crc=crc_rows(k_sel), v_crc=crc_rows(v_sel),
                            sample_rows=sample.astype(np.int32), k_rows=k_sel[:, sample], v_rows=v_sel[:, sample],
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        tally["cases"][name] = dict(S=S, layer=layer, bits=list(bits), kept=kept, ratio=ratio,
                                    oracle_vs_reference=dict(selection_flips=sel_flips, label_flips=lab_flips,
                                                             outputs_identical=same_out,
                                                             score_max_ulp=int(ud.max()), scores_differing=int((ud > 0).sum())),
                                    class_boundary_margin_ulps=margins, nearest_label_threshold_ulps=thr_ulps,
                                    reference_seconds=round(t_ref, 2))
        print(name, json.dumps(tally["cases"][name]), flush=True)
    if only:                                                   # merge into the existing tally, keep its survey
        path = os.path.join(GOLD, "PINNING_BASELINE.json")
        old = json.load(open(path))
        old["cases"].update(tally["cases"])
        with open(path, "w") as f:
            json.dump(old, f, indent=1)
        return
    # ---- flip survey: many seeded layers where the budget BINDS (tally only, no fixtures)
    n_survey = int(sys.argv[sys.argv.index("--survey") + 1]) if "--survey" in sys.argv else 48
    sv = dict(cases=0, selection_flips=0, cases_with_flips=0, label_flips=0, outputs_identical=0, score_max_ulp=0,
              min_class_boundary_margin_ulps=None, margin_histogram_ulps={"<=2": 0, "3-16": 0, "17-256": 0, ">256": 0},
              tokens=0, kept=0)
    srng = np.random.default_rng(2025)
    for t in range(n_survey):
        S = int(srng.choice([2048, 4096]))
        bits = [(8, 8, 8), (4, 8, 16), (2, 4, 8), (4, 4, 4)][t % 4]
        layer = int(srng.integers(9, 32))                      # middle / later bands: ratio 0.6 / 0.4
        K, V, A, P = baseline_inputs(S, 5000 + t)
        cfg = CompressionConfig(num_hidden_layers=L, num_attention_heads=H, hidden_size=HD,
                                high_precision_bits=bits[2], medium_precision_bits=bits[1], low_precision_bits=bits[0])
        comp = RealTimePrefillCompressor(cfg)
        with stable_argsort():
            k_sel, v_sel, info = comp.compress_layer_kv_cache(torch.from_numpy(K), torch.from_numpy(V), torch.from_numpy(A),
                                                              torch.zeros(1, S, dtype=torch.long), layer)
        scores = comp.importance_tracker.layer_scores[layer].numpy()
        labels = info["quantization_info"]["bit_assignments"].astype(np.uint8)
        mask = info["propagation_info"]["selection_mask"].numpy().astype(np.uint8)
        o = oracle.compress_layer(K, V, oracle.F32, A, oracle.F32, P, cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer],
                                  cfg.theta_h, cfg.theta_m, bits, info["propagation_info"]["propagation_ratio"])
        flips = int((mask != o["mask"]).sum())
        sv["cases"] += 1
        sv["selection_flips"] += flips
        sv["cases_with_flips"] += int(flips > 0)
        sv["label_flips"] += int((labels != o["labels"]).sum())
        sv["outputs_identical"] += int(flips == 0 and np.array_equal(k_sel.numpy(), o["k_sel"]) and np.array_equal(v_sel.numpy(), o["v_sel"]))
        sv["score_max_ulp"] = max(sv["score_max_ulp"], int(ulp_diff(scores, o["scores"]).max()))
        sv["tokens"] += S
        sv["kept"] += int(mask.sum())
        for c, m in boundary_margins(scores[0], labels[0], mask[0]).items():
            cur = sv["min_class_boundary_margin_ulps"]
            sv["min_class_boundary_margin_ulps"] = m if cur is None else min(cur, m)
            key = "<=2" if m <= 2 else "3-16" if m <= 16 else "17-256" if m <= 256 else ">256"
            sv["margin_histogram_ulps"][key] += 1
        print("survey", t, S, bits, layer, "flips", flips, flush=True)
    tally["flip_survey"] = sv
    with open(os.path.join(GOLD, "PINNING_BASELINE.json"), "w") as f:
        json.dump(tally, f, indent=1)
    print(json.dumps(sv, indent=1))


if __name__ == "__main__":
    main()
