"""Parity at BASELINE sizes against the REAL reference (tests/golden/baseline_*.npz, written by
oracle/pin_baseline_sizes.py from the unmodified reference): configs[0] [1,2048,4096] layers 0/15/31 with bits 8/4/2
and 16/8/4, one [1,4096,4096] layer and one [1,32768,4096] layer (the metric's 32k shape) per bit setting, and configs[3]
(Llama-2-13B shapes: 40 heads, width 5120, 16384 tokens) with each bit setting of its sweep (8/8/8, 4/4/4, 2/2/2, 8/4/2), and one
rank's share of configs[4] (8 sequences of 4096 tokens in one call: rows keep different numbers of tokens, zero padded), and the
configs[0] shape with K, V and the attention slice in bf16 and in f16 (every reference operation then runs in that dtype).

The inputs are regenerated from integer draws with exact elementwise arithmetic (bit-identical on every machine,
verified by CRC32 against the fixture); the reference's outputs travel as a digest: scores, labels, mask, a CRC32 of
every kept K/V row and 64 sampled rows.  Bars: selection mask, labels and every output row bit-exact; scores within
8 ulp (the H*P-term prompt-mass reduction order is ATen's; measured <= 2 ulp, PINNING_BASELINE.json)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN
from oracle.pin_baseline_sizes import CASES, as_dtype, batched_inputs, crc_all, crc_rows, ulp_diff

NAMES = [c[0] for c in CASES]


def load(name):
    z = np.load(os.path.join(GOLDEN, f"baseline_{name}.npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(bytes(z["meta"]).decode())
    return d


def inputs_for(meta):
    K, V, A, P = batched_inputs(meta["S"], meta["seed"], meta["H"], meta["HD"], meta.get("B", 1))
    assert P == meta["P"]
    name = meta.get("dtype", "f32")
    (K, dt), (V, _), (A, adt) = as_dtype(K, name), as_dtype(V, name), as_dtype(A, name)   # half types: uint16 bit patterns
    got = dict(K=crc_all(K), V=crc_all(V), A=crc_all(A))
    assert got == meta["input_crc"], "the regenerated inputs differ from the ones the reference was run on"
    return K, V, A, P, dt, adt


def check_against_digest(g, scores, labels, mask, k_sel, v_sel):
    m = g["meta"]
    np.testing.assert_array_equal(labels.astype(np.uint8), g["labels"])
    np.testing.assert_array_equal(mask.astype(np.uint8), g["mask"])
    assert int(ulp_diff(np.ascontiguousarray(scores, dtype=np.float32), g["scores"]).max()) <= 8
    assert list(k_sel.shape) == m["compressed_shape"] and list(v_sel.shape) == m["compressed_shape"]
    np.testing.assert_array_equal(crc_rows(k_sel), g["k_crc"])
    np.testing.assert_array_equal(crc_rows(v_sel), g["v_crc"])
    rows = g["sample_rows"]
    np.testing.assert_array_equal(k_sel[:, rows], g["k_rows"])
    np.testing.assert_array_equal(v_sel[:, rows], g["v_rows"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_reference_at_baseline_sizes(name):
    """CPU: the oracle restatement against the reference digest (this is what pins the oracle at these sizes)."""
    g = load(name)
    m = g["meta"]
    K, V, A, P, dt, adt = inputs_for(m)
    o = oracle.compress_layer(K, V, dt, A, adt, P, m["alpha"], m["beta"], m["gamma"], m["layer_weight"],
                              m["theta_h"], m["theta_m"], tuple(m["bits"]), m["ratio"])
    check_against_digest(g, o["scores"], o["labels"], o["mask"], o["k_sel"], o["v_sel"])


def test_flip_survey_is_recorded():
    """The tally of oracle-vs-reference selection flips where the budget binds travels with the fixtures."""
    t = json.load(open(os.path.join(GOLDEN, "PINNING_BASELINE.json")))
    assert set(NAMES) <= set(t["cases"])
    for c in t["cases"].values():
        assert c["oracle_vs_reference"]["selection_flips"] == 0 and c["oracle_vs_reference"]["outputs_identical"]
    sv = t["flip_survey"]
    assert sv["cases"] >= 40 and sv["selection_flips"] == 0 and sv["outputs_identical"] == sv["cases"]
    lg = t["flip_survey_large"]                                # 16384 / 32768 tokens, budget binding in most
    assert lg["cases"] >= 12 and lg["selection_flips"] == 0 and lg["label_flips"] == 0 and lg["outputs_identical"] == lg["cases"]
    assert lg["score_max_ulp"] <= 8 and lg["min_class_boundary_margin_ulps"] > lg["score_max_ulp"]


@pytest.mark.gpu
@pytest.mark.parametrize("on_host", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_equals_reference_at_baseline_sizes(name, on_host):
    """GPU: RealTimePrefillCompressor.compress_layer_kv_cache (CUDA tensors, and host tensors through
    rtkv_compress_layer_host) against the reference digest."""
    import rtkv
    g = load(name)
    m = g["meta"]
    if on_host and m["S"] > 4096:
        pytest.skip("host-buffer path is covered at 2048/4096 tokens")
    K, V, A, P, dt, adt = inputs_for(m)
    bits = m["bits"]
    cfg = rtkv.CompressionConfig(num_hidden_layers=m["L"], num_attention_heads=m["H"], hidden_size=m["HD"], high_precision_bits=bits[2],
                                 medium_precision_bits=bits[1], low_precision_bits=bits[0])
    comp = rtkv.RealTimePrefillCompressor(cfg)
    to = (lambda a, d: oracle.to_torch(a, d)) if on_host else (lambda a, d: oracle.to_torch(a, d).cuda())
    ids = torch.zeros(m.get("B", 1), m["S"], dtype=torch.long)
    k_sel, v_sel, info = comp.compress_layer_kv_cache(to(K, dt), to(V, dt), to(A, adt), ids if on_host else ids.cuda(), m["layer"])
    scores = comp.importance_tracker.layer_scores[m["layer"]].numpy()
    labels = info["quantization_info"]["bit_assignments"]
    mask = info["propagation_info"]["selection_mask"].cpu().numpy()
    check_against_digest(g, scores, labels, mask, oracle.from_torch(k_sel)[0], oracle.from_torch(v_sel)[0])
    assert info["precision_stats"] == m["precision_stats"]
    for k, v in m["importance_stats"].items():
        assert info["importance_stats"][k] == pytest.approx(v, rel=1e-5, abs=1e-7)
