"""CPU: the oracle (oracle/rtkv_oracle.c) against the golden fixtures that the real
reference produced (oracle/pin_against_reference.py)."""
import numpy as np
import pytest

import oracle
from conftest import LAYER_CASES, STAGE_CASES, load_golden


def ulp(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_matches_reference(name):
    g = load_golden(name)
    m = g["meta"]
    o = oracle.compress_layer(g["K"], g["V"], m["dt"], g["A"], m["dt"], m["P"], m["alpha"], m["beta"],
                              m["gamma"], m["layer_weight"], m["theta_h"], m["theta_m"], m["bits"], m["ratio"])
    # prompt-mass reduction order is ATen's in the reference: a few ulp
    assert ulp(o["scores"], g["scores"]).max() <= 8
    np.testing.assert_array_equal(o["labels"], g["labels"])
    np.testing.assert_array_equal(o["mask"], g["mask"])          # bit-exact retained index set
    assert o["max_sel"] == m["compressed_shape"][1]
    np.testing.assert_array_equal(o["k_sel"], g["k_sel"])        # bit-exact fake-quant values
    np.testing.assert_array_equal(o["v_sel"], g["v_sel"])


@pytest.mark.parametrize("name", LAYER_CASES)
def test_stages_exact_given_reference_scores(name):
    g = load_golden(name)
    m = g["meta"]
    labels, counts = oracle.assign_precision(g["scores"], m["theta_h"], m["theta_m"])
    np.testing.assert_array_equal(labels, g["labels"])
    ps = m["precision_stats"]
    assert [int(counts[0]), int(counts[1]), int(counts[2])] == [ps["low_count"], ps["medium_count"], ps["high_count"]]
    mask, cnt, cost, ssum, cc = oracle.select_tokens(g["scores"], labels, m["bits"], m["S"] * m["ratio"])
    if m["name"].endswith("fallback"):
        assert cnt.max() == 0
        k = max(1, int(m["S"] * 0.1))
        np.testing.assert_array_equal(oracle.topk_mask(g["scores"], k), g["mask"])
    else:
        np.testing.assert_array_equal(mask, g["mask"])
        assert list(cnt[cnt > 0]) == m["selected_counts"]
        np.testing.assert_allclose(cost[cnt > 0] / (m["S"] * m["ratio"]), m["budget_utilization"], rtol=1e-12)
        np.testing.assert_allclose(ssum[cnt > 0] / cnt[cnt > 0], m["avg_importance"], rtol=1e-5)
        cd = m["cost_distribution"]
        assert [int(cc[:, 0].sum()), int(cc[:, 1].sum()), int(cc[:, 2].sum())] == [cd["low"], cd["medium"], cd["high"]]


@pytest.mark.parametrize("name", STAGE_CASES)
def test_stage_fixtures(name):
    g = load_golden(name)
    m = g["meta"]
    labels, _ = oracle.assign_precision(g["scores"], m["theta_h"], m["theta_m"])
    np.testing.assert_array_equal(labels, g["labels"])
    kq, _, _ = oracle.fake_quant_rows(g["K"], m["dt"], labels, m["bits"])
    vq, _, _ = oracle.fake_quant_rows(g["V"], m["dt"], labels, m["bits"])
    np.testing.assert_array_equal(kq, g["k_quant"])
    np.testing.assert_array_equal(vq, g["v_quant"])
    budget = m["S"] * m["ratio"]
    for lab, want in ((labels, g["mask_mono"]), (g["rand_labels"], g["mask_rand"])):
        mask, cnt, _, _, _ = oracle.select_tokens(g["scores"], lab, m["bits"], budget)
        np.testing.assert_array_equal(mask, want)
    mask, cnt, _, _, _ = oracle.select_tokens(g["scores"], g["rand_labels"], m["bits"], budget)
    ms = int(cnt.max())
    np.testing.assert_array_equal(oracle.gather_rows(g["K"], m["dt"], mask, ms), g["k_sel"])
    np.testing.assert_array_equal(oracle.gather_rows(g["V"], m["dt"], mask, ms), g["v_sel"])


def test_constant_row_quirk():
    """scale=1, zp=0 when max==min: a row of -2.0 dequantises to 0.0 (dynamic_quantization.py:83-86)."""
    x = np.full((1, 1, 16), -2.0, np.float32)
    out, codes, qp = oracle.fake_quant_rows(x, oracle.F32, np.zeros((1, 1), np.uint8), (2, 4, 8), True, True)
    assert (out == 0).all() and qp[0, 0] == 1.0 and qp[0, 1] == 0.0
    x[:] = 2.5   # rint half-to-even: 2.5 -> 2
    out, _, _ = oracle.fake_quant_rows(x, oracle.F32, np.full((1, 1), 2, np.uint8), (2, 4, 8))
    assert (out == 2.0).all()


def test_budget_is_double_arithmetic():
    """4096*0.8 = 3276.8: all-MID 4-bit tokens (cost 0.5) -> 6553 fit; 8-bit (1.0) -> 3276 (+ cheaper ones)."""
    S = 4096
    scores = np.linspace(1, 0, S, dtype=np.float32)[None]
    lab = np.ones((1, S), np.uint8)
    lab[0, -10:] = 0
    mask, cnt, cost, _, cc = oracle.select_tokens(scores, lab, (2, 8, 8), S * 0.8)
    assert cc[0, 1] == 3276 and cc[0, 0] == 3 and cost[0] == 3276.75   # 3 LOW tokens (0.25 each) still fit


def test_ties_lowest_index_first():
    scores = np.zeros((1, 40), np.float32)
    lab = np.full((1, 40), 2, np.uint8)
    mask, cnt, _, _, _ = oracle.select_tokens(scores, lab, (8, 8, 8), 10.0)
    assert cnt[0] == 10 and mask[0, :10].all() and not mask[0, 10:].any()


def test_position_bias_properties():
    pb = oracle.position_bias(50)        # tests/test_importance_scoring.py:80-102
    assert pb[0] == 0 and abs(pb[-1] - 1) < 1e-6 and (np.diff(pb) >= 0).all()
    assert (oracle.position_bias(1) == 0).all()
    assert oracle.context_relevance(20, 4) == np.float32(0.2)


def test_extended_quant_restatement_properties():
    """orc_fake_quant_ex (project-defined extras): the default mode is the reference restatement, the outlier
    split does not depend on the order of the channels inside a group, per-channel scales do not depend on
    the order of the rows inside a token group."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((24, 256)).astype(np.float32)
    x[rng.integers(0, 24, 30), rng.integers(0, 256, 30)] *= 40.0
    lab = rng.integers(0, 3, 24).astype(np.uint8)
    bits = (2, 4, 8)
    ref, _, _ = oracle.fake_quant_rows(x, oracle.F32, lab, bits)
    r0 = oracle.fake_quant_ex(x, oracle.F32, lab, bits)
    np.testing.assert_array_equal(r0["out"].view(np.uint32), ref.view(np.uint32))
    assert r0["omask"].sum() == 0
    # outliers: permuting the channels of every 64-wide group permutes the result
    r1 = oracle.fake_quant_ex(x, oracle.F32, lab, bits, mode=1, group=64, tau=5.0)
    assert 0 < r1["omask"].sum() < 200
    perm = np.concatenate([g * 64 + rng.permutation(64) for g in range(4)])
    r2 = oracle.fake_quant_ex(x[:, perm], oracle.F32, lab, bits, mode=1, group=64, tau=5.0)
    np.testing.assert_array_equal(r2["out"].view(np.uint32), r1["out"][:, perm].view(np.uint32))
    np.testing.assert_array_equal(r2["omask"], r1["omask"][:, perm])
    kept = r1["omask"] == 1
    np.testing.assert_array_equal(r1["out"][kept], x[kept])                  # outliers are kept exactly
    assert np.abs(r1["out"] - x)[~kept].mean() < np.abs(ref - x).mean()      # and the rest quantises finer
    # per-channel mode: permuting the rows inside a token group permutes the result
    rows = np.concatenate([g * 8 + rng.permutation(8) for g in range(3)])
    c1 = oracle.fake_quant_ex(x, oracle.F32, lab, bits, mode=2, tgroup=8)
    c2 = oracle.fake_quant_ex(x[rows], oracle.F32, lab[rows], bits, mode=2, tgroup=8)
    np.testing.assert_array_equal(c2["out"].view(np.uint32), c1["out"][rows].view(np.uint32))
    np.testing.assert_array_equal(c2["qparams"].view(np.uint32), c1["qparams"].view(np.uint32))
    # a cap on the split outliers: later outliers are clamped codes, not exact values
    r3 = oracle.fake_quant_ex(x, oracle.F32, lab, bits, mode=0, tau=3.0, cap=1)
    assert r3["omask"].sum(axis=1).max() == 1


def test_oracle_matches_reference_on_edge_rows():
    """tests/golden/edge_rows_f16.npz holds the REAL reference's fp16 output (oracle/pin_edge_rows.py) on the stress and
    fuzz rows the GPU tests use: wide dynamic range, zeros of both signs, denormals, a constant row that rounds to -0.0,
    non-finite rows and narrow rows whose x/scale overflows the half range (NaN in the reference)."""
    import os
    import torch
    from test_parity_gpu import fuzz_rows, stress_rows
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edge_rows_f16.npz"))
    rows = np.stack(stress_rows(np.random.default_rng(77), 256) + list(fuzz_rows(np.random.default_rng(1000), 400, 256)))
    x = oracle.from_torch(torch.from_numpy(rows[None]).to(torch.float16))[0]
    got, _, _ = oracle.fake_quant_rows(x, oracle.F16, fx["labels"][None], tuple(int(b) for b in fx["bits"]))
    g = oracle.to_torch(got[0], oracle.F16).float().numpy()
    w = oracle.to_torch(fx["want"], oracle.F16).float().numpy()
    assert np.isnan(w).sum() > 10000 and ((w == 0) & np.signbit(w)).sum() > 1000        # the fixture does hold the edge cases
    np.testing.assert_array_equal(np.isnan(g), np.isnan(w))
    ok = ~np.isnan(w)
    np.testing.assert_array_equal(g[ok], w[ok])
    np.testing.assert_array_equal(np.signbit(g[ok]), np.signbit(w[ok]))
