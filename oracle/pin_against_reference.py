#!/usr/bin/env python
"""Pin the CPU oracle against the real reference and write tests/golden/.

Runs ONLY in the build container (needs /root/reference, read-only).  It
  1. imports the unmodified reference (src.compression.*, configs.base_config),
  2. runs it on seeded inputs and stores inputs + outputs as small fixtures
     under tests/golden/ (these travel to the GPU box; the reference does not),
  3. compares the C oracle with the reference stage by stage over a few hundred
     random cases and writes the tally to tests/golden/PINNING.json.

Usage:  python oracle/pin_against_reference.py [--trials N]
"""
import argparse
import contextlib
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import oracle  # noqa: E402
from configs.base_config import CompressionConfig  # noqa: E402  (reference)
from src.compression.dynamic_quantization import DynamicPrecisionQuantizer  # noqa: E402
from src.compression.selective_propagation import SelectiveTokenPropagator  # noqa: E402
from src.compression.token_importance import PromptGuidedImportanceScorer  # noqa: E402
from src.compression.unified_compressor import RealTimePrefillCompressor  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TDT = {oracle.F32: torch.float32, oracle.F16: torch.float16, oracle.BF16: torch.bfloat16}


@contextlib.contextmanager
def stable_argsort():
    """The reference's argsort is unstable (tie order is noise for n>16, SURVEY.md
    App. A).  The contract is lowest-index-first; this makes the reference obey it."""
    orig = torch.argsort

    def patched(x, dim=-1, descending=False, stable=False):
        return orig(x, dim=dim, descending=descending, stable=True)

    torch.argsort = patched
    try:
        yield
    finally:
        torch.argsort = orig


def causal_probs(B, H, S, dtype, gen, sharp=1.0):
    q = torch.randn(B, H, S, 16, generator=gen)
    k = torch.randn(B, H, S, 16, generator=gen)
    logits = sharp * q @ k.transpose(-1, -2) / 4.0
    logits = logits + torch.full((S, S), float("-inf")).triu(1)
    return torch.softmax(logits, dim=-1, dtype=torch.float32).to(dtype)


def run_reference_layer(cfg, K, V, A, layer_idx):
    comp = RealTimePrefillCompressor(cfg)
    ids = torch.zeros(K.shape[0], K.shape[1], dtype=torch.long)
    with stable_argsort():
        k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, layer_idx)
    scores = comp.importance_tracker.layer_scores[layer_idx]
    return k_sel, v_sel, info, scores


def t2n(t):
    a, _ = oracle.from_torch(t)
    return a


def golden_layer(name, seed, B, H, S, HD, dt, bits, L, layer_idx, **cfg_kw):
    gen = torch.Generator().manual_seed(seed)
    tdt = TDT[dt]
    A = causal_probs(B, H, S, tdt, gen)
    K = torch.randn(B, S, HD, generator=gen).to(tdt)
    V = (torch.randn(B, S, HD, generator=gen) * 0.5 + 0.1).to(tdt)
    cfg = CompressionConfig(num_hidden_layers=L, num_attention_heads=H, hidden_size=HD,
                            high_precision_bits=bits[2], medium_precision_bits=bits[1],
                            low_precision_bits=bits[0], **cfg_kw)
    k_sel, v_sel, info, scores = run_reference_layer(cfg, K, V, A, layer_idx)
    P = max(1, min(S // 5, 128))
    # the reference only reads the first P key columns: the narrow slice is bit-identical
    k2, v2, info2, scores2 = run_reference_layer(cfg, K, V, A[..., :P].contiguous(), layer_idx)
    assert torch.equal(scores, scores2) and torch.equal(k_sel, k2) and torch.equal(v_sel, v2)
    pinfo = info["propagation_info"]
    st = pinfo["selection_stats"]
    meta = dict(
        name=name, seed=seed, B=B, H=H, S=S, HD=HD, P=P, dt=dt, bits=list(bits), L=L,
        layer_idx=layer_idx, alpha=cfg.alpha, beta=cfg.beta, gamma=cfg.gamma,
        theta_h=cfg.theta_h, theta_m=cfg.theta_m, layer_weight=cfg.layer_weights[layer_idx],
        ratio=pinfo["propagation_ratio"], early=cfg.early_layer_ratio,
        middle=cfg.middle_layer_ratio, later=cfg.later_layer_ratio,
        compressed_shape=list(info["compressed_shape"]),
        compression_ratio=info["compression_ratio"], importance_stats=info["importance_stats"],
        precision_stats=info["precision_stats"], selected_counts=st["selected_counts"],
        budget_utilization=st["budget_utilization"], avg_importance=st["avg_importance"],
        cost_distribution=st["cost_distribution"],
    )
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        K=t2n(K), V=t2n(V), A=t2n(A[..., :P].contiguous()),
        k_sel=t2n(k_sel), v_sel=t2n(v_sel), scores=scores.numpy(),
        labels=info["quantization_info"]["bit_assignments"].astype(np.uint8),
        mask=pinfo["selection_mask"].numpy().astype(np.uint8),
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    return meta


def golden_stage(name, seed, B, S, HD, dt, bits, ratio_layer, L=32):
    """Stage fixtures in the style of the reference's unit tests: uniform scores
    (tests/test_quantization.py:30) and *random* labels (tests/test_compression.py:87)."""
    gen = torch.Generator().manual_seed(seed)
    tdt = TDT[dt]
    cfg = CompressionConfig(num_hidden_layers=L, high_precision_bits=bits[2],
                            medium_precision_bits=bits[1], low_precision_bits=bits[0])
    scores = torch.rand(B, S, generator=gen)
    K = torch.randn(B, S, HD, generator=gen).to(tdt)
    V = torch.randn(B, S, HD, generator=gen).to(tdt)
    K[0, 1, :] = -2.0  # constant row quirk (dynamic_quantization.py:83-86)
    V[0, 2, :] = 0.75
    quant = DynamicPrecisionQuantizer(cfg)
    prop = SelectiveTokenPropagator(cfg)
    labels, stats = quant.assign_precision_levels(scores)
    kq, vq, _ = quant.apply_mixed_precision_quantization(K, V, labels)
    rand_labels = torch.randint(0, 3, (B, S), generator=gen)
    with stable_argsort():
        mask_mono, _ = prop.select_tokens_with_budget(scores, labels, prop.get_layer_propagation_ratio(ratio_layer), ratio_layer)
        mask_rand, _ = prop.select_tokens_with_budget(scores, rand_labels, prop.get_layer_propagation_ratio(ratio_layer), ratio_layer)
        ks, vs, ss, ls, pinfo = prop.apply_token_selection(K, V, scores, rand_labels, ratio_layer)
    meta = dict(name=name, seed=seed, B=B, S=S, HD=HD, dt=dt, bits=list(bits), L=L,
                theta_h=cfg.theta_h, theta_m=cfg.theta_m,
                ratio=prop.get_layer_propagation_ratio(ratio_layer), ratio_layer=ratio_layer,
                precision_stats=stats)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        scores=scores.numpy(), K=t2n(K), V=t2n(V), labels=labels.numpy().astype(np.uint8),
        k_quant=t2n(kq), v_quant=t2n(vq), rand_labels=rand_labels.numpy().astype(np.uint8),
        mask_mono=mask_mono.numpy().astype(np.uint8), mask_rand=mask_rand.numpy().astype(np.uint8),
        k_sel=t2n(ks), v_sel=t2n(vs), sel_scores=ss.numpy(), sel_labels=ls.numpy().astype(np.uint8),
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    return meta


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


def pin_trials(trials):
    """Oracle vs reference, stage by stage, on random cases."""
    rng = np.random.default_rng(1234)
    tally = dict(trials=0, mass_max_ulp=0, mass_max_rel=0.0, score_max_ulp=0, score_neq=0, score_elems=0,
                 posbias_neq=0, posbias_elems=0,
                 scores_given_mass_neq=0, labels_neq=0, quant_neq=0, quant_elems=0,
                 select_neq=0, select_cases=0, layer_mask_neq=0, layer_out_neq=0, layer_cases=0,
                 layer_cases_scores_equal=0)
    for t in range(trials):
        dt = [oracle.F32, oracle.F32, oracle.BF16, oracle.F16][t % 4]
        tdt = TDT[dt]
        B = int(rng.integers(1, 3))
        H = int(rng.choice([4, 8, 32, 40]))
        S = int(rng.choice([2, 5, 7, 16, 17, 40, 64, 129, 200, 333, 700]))
        HD = int(rng.choice([8, 64, 96, 128, 257]))
        bits = [(2, 4, 8), (4, 8, 16), (8, 8, 8), (4, 4, 4), (2, 2, 2), (2, 4, 8)][int(rng.integers(0, 6))]
        L = int(rng.choice([4, 32, 40]))
        layer = int(rng.integers(0, L))
        if dt == oracle.F16 and max(bits) >= 16:
            # reference raises: torch.clamp(x, 0, 65535) overflows fp16 (dynamic_quantization.py:121)
            bits = (2, 4, 8)
        gen = torch.Generator().manual_seed(10_000 + t)
        A = causal_probs(B, H, S, tdt, gen, sharp=float(rng.choice([0.2, 1.0, 4.0])))
        K = (torch.randn(B, S, HD, generator=gen) * float(rng.choice([0.02, 1.0, 30.0]))).to(tdt)
        V = torch.randn(B, S, HD, generator=gen).to(tdt)
        cfg = CompressionConfig(num_hidden_layers=L, num_attention_heads=H,
                                high_precision_bits=bits[2], medium_precision_bits=bits[1],
                                low_precision_bits=bits[0])
        scorer = PromptGuidedImportanceScorer(cfg)
        quant = DynamicPrecisionQuantizer(cfg)
        prop = SelectiveTokenPropagator(cfg)
        P = max(1, min(S // 5, 128))
        pidx = torch.arange(P)
        # --- stage: prompt mass (reduction order is ATen's: tolerance) ---
        ref_mass = scorer.compute_attention_aggregation(A, pidx, layer)
        An, adt = oracle.from_torch(A)
        orc_mass = oracle.prompt_mass(An, adt, P)
        rm = ref_mass.float().numpy()
        tally["mass_max_ulp"] = max(tally["mass_max_ulp"], int(ulp_diff(rm, orc_mass).max()))
        tally["mass_max_rel"] = max(tally["mass_max_rel"], float((np.abs(rm - orc_mass) / np.maximum(np.abs(rm), 1e-30)).max()))
        # --- stage: position bias ---
        ref_pb = scorer.compute_position_bias(S, torch.device("cpu")).numpy()
        tally["posbias_neq"] += int((ref_pb != oracle.position_bias(S)).sum())
        tally["posbias_elems"] += S
        # --- stage: scores given the *reference's* mass, position bias aside ---
        ref_scores = scorer.compute_importance_scores(A, pidx, layer).float().numpy()
        orc_scores = oracle.importance_scores(rm, P, cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer], adt)
        tally["score_max_ulp"] = max(tally["score_max_ulp"], int(ulp_diff(ref_scores, orc_scores).max()))
        tally["score_neq"] += int((ref_scores != orc_scores).sum())
        tally["score_elems"] += ref_scores.size
        # exactness of everything but logf: rebuild scores with the reference's own position bias
        nrm = oracle.normalize(rm, adt)
        t1 = nrm.copy()
        if adt == oracle.F32:
            t1 = (np.float32(cfg.alpha) * nrm) * np.float32(cfg.layer_weights[layer])
        else:
            t1 = torch.from_numpy(nrm).to(tdt)
            t1 = ((t1 * cfg.alpha) * cfg.layer_weights[layer]).float().numpy()
        rebuilt = (t1 + np.float32(cfg.beta) * ref_pb[None, :]) + np.float32(cfg.gamma) * np.float32(oracle.context_relevance(S, P))
        tally["scores_given_mass_neq"] += int((rebuilt.astype(np.float32) != ref_scores).sum())
        # --- stage: labels given the reference's scores ---
        ref_scores_t = torch.from_numpy(ref_scores)
        ref_labels, _ = quant.assign_precision_levels(ref_scores_t)
        orc_labels, _ = oracle.assign_precision(ref_scores, cfg.theta_h, cfg.theta_m)
        tally["labels_neq"] += int((ref_labels.numpy() != orc_labels).sum())
        # --- stage: fake-quant given labels (random labels to exercise all widths) ---
        lab = torch.from_numpy(rng.integers(0, 3, (B, S)))
        if S > 1:
            K[0, 1, :] = K[0, 1, 0]  # constant row
        kq, vq, _ = quant.apply_mixed_precision_quantization(K, V, lab)
        Kn, kdt = oracle.from_torch(K)
        Vn, _ = oracle.from_torch(V)
        ok, _, _ = oracle.fake_quant_rows(Kn, kdt, lab.numpy().astype(np.uint8), bits)
        ov, _, _ = oracle.fake_quant_rows(Vn, kdt, lab.numpy().astype(np.uint8), bits)
        tally["quant_neq"] += int((t2n(kq) != ok).sum() + (t2n(vq) != ov).sum())
        tally["quant_elems"] += 2 * ok.size
        # --- stage: selection given scores + labels (monotone and random) ---
        for labs in (ref_labels, lab):
            ratio = prop.get_layer_propagation_ratio(layer)
            with stable_argsort():
                ref_mask, _ = prop.select_tokens_with_budget(ref_scores_t, labs, ratio, layer)
            om, _, _, _, _ = oracle.select_tokens(ref_scores, labs.numpy().astype(np.uint8), bits, S * ratio)
            tally["select_neq"] += int((ref_mask.numpy().astype(np.uint8) != om).sum())
            tally["select_cases"] += 1
        # --- whole layer ---
        k_sel, v_sel, info, sc = run_reference_layer(cfg, K, V, A, layer)
        o = oracle.compress_layer(Kn, Vn, kdt, An, adt, P, cfg.alpha, cfg.beta, cfg.gamma,
                                  cfg.layer_weights[layer], cfg.theta_h, cfg.theta_m, bits,
                                  prop.get_layer_propagation_ratio(layer))
        same_mask = np.array_equal(info["propagation_info"]["selection_mask"].numpy().astype(np.uint8), o["mask"])
        tally["layer_cases"] += 1
        tally["layer_mask_neq"] += int(not same_mask)
        if same_mask:
            tally["layer_out_neq"] += int(not (np.array_equal(t2n(k_sel), o["k_sel"]) and np.array_equal(t2n(v_sel), o["v_sel"])))
        tally["layer_cases_scores_equal"] += int(np.array_equal(sc.numpy(), o["scores"]))
        tally["trials"] += 1
    return tally


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=120)
    args = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLD, exist_ok=True)
    metas = [
        golden_layer("layer_f32_s64", 1, 1, 8, 64, 64, oracle.F32, (2, 4, 8), 4, 0),
        golden_layer("layer_f32_b2_s96", 2, 2, 8, 96, 128, oracle.F32, (4, 8, 16), 32, 12),
        golden_layer("layer_f32_s256_l31", 3, 1, 32, 256, 128, oracle.F32, (2, 4, 8), 32, 31),
        golden_layer("layer_f32_s700_888", 4, 1, 4, 700, 32, oracle.F32, (8, 8, 8), 32, 3),
        golden_layer("layer_bf16_s128", 5, 1, 8, 128, 128, oracle.BF16, (2, 4, 8), 32, 5),
        golden_layer("layer_f16_s128", 6, 2, 8, 128, 64, oracle.F16, (2, 4, 8), 32, 20),
        golden_layer("layer_f32_tiny_fallback", 7, 1, 4, 2, 8, oracle.F32, (4, 8, 16), 4, 3),
        golden_stage("stage_f32_8_4_2", 11, 2, 48, 64, oracle.F32, (2, 4, 8), 0),
        golden_stage("stage_f32_16_8_4", 12, 2, 80, 96, oracle.F32, (4, 8, 16), 25),
        golden_stage("stage_bf16_8_4_2", 13, 1, 64, 64, oracle.BF16, (2, 4, 8), 10),
        golden_stage("stage_f16_8_4_2", 14, 1, 64, 64, oracle.F16, (2, 4, 8), 30),
    ]
    tally = pin_trials(args.trials)
    tally["torch"] = torch.__version__
    tally["golden"] = [m["name"] for m in metas]
    with open(os.path.join(GOLD, "PINNING.json"), "w") as f:
        json.dump(tally, f, indent=1)
    print(json.dumps(tally, indent=1))


if __name__ == "__main__":
    main()
