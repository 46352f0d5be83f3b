"""Extended scale granularities and the outlier split (north-star extras, no reference behaviour:
SURVEY.md §8c "parity unpinned by the reference").  The CUDA kernels are compared bit for bit with the
project's own CPU restatement (oracle.fake_quant_ex); the default options must reproduce the
reference-semantics kernel exactly."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
TDT = {oracle.F32: torch.float32, oracle.BF16: torch.bfloat16, oracle.F16: torch.float16}
MODE = {"token": 0, "token_group": 1, "channel": 2}


@pytest.fixture(scope="module")
def rt(built_lib):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv
    from rtkv import ops
    return rtkv, ops


def same(got, want, dt):
    """Bit-exact (values and zero signs; NaN where the oracle has NaN, payloads not compared)."""
    from test_parity_gpu import assert_same
    assert_same(got.reshape(want.shape), want, dt)


def same_values(got, want, dt):
    """Unpacked rows: equal values; an integer code cannot carry the sign of a zero, so a -0.0 of the dense
    path (constant group in (-0.5, 0)) comes back as +0.0."""
    g = got.detach().float().cpu().numpy().reshape(want.shape)
    np.testing.assert_array_equal(g, oracle.to_torch(np.ascontiguousarray(want), dt).float().numpy())


def same_f32(got, want):
    g, w = got.detach().cpu().numpy().reshape(want.shape), want
    np.testing.assert_array_equal(g, w)
    ok = ~np.isnan(w)
    np.testing.assert_array_equal(np.signbit(g)[ok], np.signbit(w)[ok])


def make_rows(rng, B, S, n, dt, heavy=True):
    x = rng.standard_normal((B, S, n)).astype(np.float32) * rng.choice([0.05, 1.0, 20.0])
    if heavy:                                        # a few large entries per row: the outliers
        for b in range(B):
            for s in range(S):
                k = rng.integers(0, 4)
                x[b, s, rng.integers(0, n, k)] *= 25.0
    x[0, 1] = -2.0                                   # constant row
    x[0, 2] = 0.0
    if n > 3:
        x[0, 3, 1] = float("nan")
        x[0, 4, 2] = float("inf")
    return torch.from_numpy(x).to(TDT[dt])


def selection(ops, rng, B, S, bits):
    scores = rng.random((B, S), dtype=np.float32)
    scores[B - 1] *= 0.3
    labels, _ = oracle.assign_precision(scores, 0.7, 0.3)
    from test_parity_gpu import run_select
    sel, layer, rows = run_select(ops, scores, labels, bits, S * 0.3)
    return sel, labels, int(layer["max_selected"]), rows


def oracle_dense(x, dt, labels, mask, ms, bits, mode, group, tgroup, tau):
    """Per batch row: kept rows (ascending) -> oracle -> zero padded [B, ms, n]."""
    xn, _ = oracle.from_torch(x)
    B, S, n = xn.shape
    out = np.zeros((B, ms, n), xn.dtype)
    counts = np.zeros((B, ms), np.int32)
    for b in range(B):
        keep = np.nonzero(mask[b])[0]
        if keep.size == 0:
            continue
        r = oracle.fake_quant_ex(xn[b][keep], dt, labels[b][keep], bits, mode=mode, group=group, tgroup=tgroup, tau=tau)
        out[b, :keep.size] = r["out"]
        counts[b, :keep.size] = r["omask"].sum(axis=1)
    return out, counts


def test_default_options_are_the_reference_kernel(rt):
    _, ops = rt
    rng = np.random.default_rng(0)
    B, S, n = 2, 40, 512
    x, v = make_rows(rng, B, S, n, oracle.F32), make_rows(rng, B, S, n, oracle.F32)
    lab = torch.from_numpy(rng.integers(0, 3, (B, S)).astype(np.uint8)).cuda()
    a = ops.quant_gather(x.cuda(), v.cuda(), None, lab, (2, 4, 8), S)
    for opt in (None, ops.quant_options()):
        b = ops.quant_gather_ex(x.cuda(), v.cuda(), None, lab, (2, 4, 8), S, opt, want_counts=True)
        assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32))
        assert torch.equal(a[1].view(torch.int32), b[1].view(torch.int32))
        assert int(b[2].abs().sum()) == 0


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("n,group", [(512, 128), (4096, 128), (4096, 32), (5120, 256), (1024, 8), (96, 48), (1000, 100),
                                     (4096, 4096)])
@pytest.mark.parametrize("tau", [0.0, 6.0])
def test_token_group_scales_and_outliers(rt, dt, n, group, tau):
    _, ops = rt
    rng = np.random.default_rng(n + group)
    B, S, bits = 2, 60, (2, 4, 8)
    K, V = make_rows(rng, B, S, n, dt), make_rows(rng, B, S, n, dt)
    sel, labels, ms, rows = selection(ops, rng, B, S, bits)
    mask = sel.mask.cpu().numpy()
    opt = ops.quant_options("token_group", "token", group_size=group, outlier_tau=tau)
    yk, yv, oc = ops.quant_gather_ex(K.cuda(), V.cuda(), sel, sel.labels, bits, ms, opt, want_counts=True)
    wk, ck = oracle_dense(K, dt, labels, mask, ms, bits, 1, group, 0, tau)
    wv, cv = oracle_dense(V, dt, labels, mask, ms, bits, 0, 0, 0, tau)
    same(yk, wk, dt)
    same(yv, wv, dt)
    np.testing.assert_array_equal(oc.cpu().numpy()[..., 0], ck)
    np.testing.assert_array_equal(oc.cpu().numpy()[..., 1], cv)
    if tau:
        assert ck.sum() > 0 and cv.sum() > 0


@pytest.fixture(params=["l2-reread", "cp.async-staged"])
def chan_variant(request, rt):
    from rtkv._lib import lib
    lib.rtkv_debug_channel_variant(1 if request.param == "cp.async-staged" else 0)
    yield request.param
    lib.rtkv_debug_channel_variant(0)


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("n,tg", [(512, 32), (4096, 32), (4096, 16), (5120, 8), (1024, 5), (100, 7), (8, 1), (2048, 20)])
def test_per_channel_keys(rt, chan_variant, dt, n, tg):
    _, ops = rt
    rng = np.random.default_rng(n + tg)
    B, S, bits = 2, 75, (2, 4, 8)
    K, V = make_rows(rng, B, S, n, dt, heavy=False), make_rows(rng, B, S, n, dt, heavy=False)
    sel, labels, ms, rows = selection(ops, rng, B, S, bits)
    mask = sel.mask.cpu().numpy()
    g = 128 if n % 128 == 0 else n
    opt = ops.quant_options("channel", "token_group", group_size=g, token_group=tg)
    yk, yv = ops.quant_gather_ex(K.cuda(), V.cuda(), sel, sel.labels, bits, ms, opt)
    wk, _ = oracle_dense(K, dt, labels, mask, ms, bits, 2, 0, tg, 0.0)
    wv, _ = oracle_dense(V, dt, labels, mask, ms, bits, 1, g, 0, 0.0)
    same(yk, wk, dt)
    same(yv, wv, dt)


def test_all_rows_without_selection_and_no_zero_pad(rt):
    _, ops = rt
    rng = np.random.default_rng(5)
    B, S, n = 1, 33, 256
    x = make_rows(rng, B, S, n, oracle.F32)
    labels = rng.integers(0, 3, (B, S)).astype(np.uint8)
    opt = ops.quant_options("channel", token_group=32)
    y, _ = ops.quant_gather_ex(x.cuda(), None, None, torch.from_numpy(labels).cuda(), (4, 4, 8), S, opt)
    r = oracle.fake_quant_ex(oracle.from_torch(x)[0][0], oracle.F32, labels[0], (4, 4, 8), mode=2, tgroup=32)
    same(y, r["out"][None], oracle.F32)


def test_bad_options_are_rejected(rt):
    _, ops = rt
    x = torch.randn(1, 4, 64, device="cuda")
    lab = torch.zeros(1, 4, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="group_size"):
        ops.quant_gather_ex(x, None, None, lab, (2, 4, 8), 4, ops.quant_options("token_group", group_size=48))
    with pytest.raises(RuntimeError, match="token_group"):
        ops.quant_gather_ex(x, None, None, lab, (2, 4, 8), 4, ops.quant_options("channel", token_group=64))
    with pytest.raises(RuntimeError, match="tau"):
        ops.quant_gather_ex(x, None, None, lab, (2, 4, 8), 4, ops.quant_options(outlier_tau=0.5))


# ---------------------------------------------------------------- packed rows with options
@pytest.mark.parametrize("dt,bits", [(oracle.F32, (2, 4, 8)), (oracle.BF16, (2, 4, 8)), (oracle.F32, (4, 8, 16)),
                                     (oracle.F16, (1, 2, 12))])
@pytest.mark.parametrize("mode,group,tau,cap", [("token", 0, 6.0, 16), ("token_group", 64, 0.0, 0),
                                                ("token_group", 128, 5.0, 3), ("token", 0, 0.0, 0),
                                                ("token_group", 24, 0.0, 0)])      # 24: not a power of two -> scalar kernel
def test_packed_token_modes(rt, dt, bits, mode, group, tau, cap):
    _, ops = rt
    rng = np.random.default_rng(bits[0] * 7 + group)
    B, S, n = 2, 48, (480 if group == 24 else 512)
    x = make_rows(rng, B, S, n, dt)
    sel, labels, ms, rows = selection(ops, rng, B, S, bits)
    mask = sel.mask.cpu().numpy()
    opt = ops.quant_options(mode, mode, group_size=group, outlier_tau=tau)
    pk = ops.quant_pack_ex(x.cuda(), sel, sel.labels, bits, ms, opt, tensor=0, outlier_cap=cap)
    xn, _ = oracle.from_torch(x)
    xf = x.float().numpy()
    rb = ops.packed_row_bytes(n, max(bits))
    G = n // group if mode == "token_group" else 1
    want_dense = np.zeros((B, ms, n), xn.dtype)
    for b in range(B):
        keep = np.nonzero(mask[b])[0]
        r = oracle.fake_quant_ex(xn[b][keep], dt, labels[b][keep], bits, mode=MODE[mode], group=group, tau=tau,
                                 cap=cap if tau else -1)
        c = keep.size
        np.testing.assert_array_equal(pk["codes"][b, :c].cpu().numpy(),
                                      oracle.pack_codes(r["codes"], labels[b][keep], bits, rb))       # integer codes
        assert int(pk["codes"][b, c:].sum()) == 0
        same_f32(pk["qparams"][b, :c], r["qparams"].reshape(c, G, 2))
        if tau:
            cnt = pk["out_cnt"][b].cpu().numpy()
            np.testing.assert_array_equal(cnt[:c], r["omask"].sum(axis=1))
            assert (cnt[c:] == 0).all() and cnt.max() <= cap
            for i in range(c):
                idx = np.nonzero(r["omask"][i])[0]
                np.testing.assert_array_equal(pk["out_idx"][b, i, :idx.size].cpu().numpy(), idx)
                np.testing.assert_array_equal(pk["out_val"][b, i, :idx.size].cpu().numpy(), xf[b][keep][i][idx])
        want_dense[b, :c] = r["out"]
    y = ops.unpack_dequant_ex(pk, TDT[dt], S, n, sel, sel.labels, bits, opt, tensor=0)
    same_values(y, want_dense, dt)                                    # unpacked == the dense fake-quant path
    if not tau or cap >= 16:
        d = ops.quant_gather_ex(x.cuda(), None, sel, sel.labels, bits, ms, opt)[0]
        assert torch.equal(torch.nan_to_num(d.float(), nan=12345.0), torch.nan_to_num(y.float(), nan=12345.0))


@pytest.mark.parametrize("dt,bits", [(oracle.F32, (2, 4, 8)), (oracle.BF16, (1, 4, 8)), (oracle.F32, (4, 8, 16)),
                                     (oracle.F32, (1, 2, 4))])          # fp32 with 1-bit codes: the scalar kernels
@pytest.mark.parametrize("n,tg", [(512, 32), (264, 5)])
def test_packed_channel_mode(rt, dt, bits, n, tg):
    _, ops = rt
    rng = np.random.default_rng(n + bits[0])
    B, S = 2, 70
    x = make_rows(rng, B, S, n, dt, heavy=False)
    sel, labels, ms, rows = selection(ops, rng, B, S, bits)
    mask = sel.mask.cpu().numpy()
    opt = ops.quant_options("channel", token_group=tg)
    pk = ops.quant_pack_ex(x.cuda(), sel, sel.labels, bits, ms, opt, tensor=0)
    xn, _ = oracle.from_torch(x)
    rb = ops.packed_row_bytes(n, max(bits))
    want_dense = np.zeros((B, ms, n), xn.dtype)
    for b in range(B):
        keep = np.nonzero(mask[b])[0]
        c = keep.size
        r = oracle.fake_quant_ex(xn[b][keep], dt, labels[b][keep], bits, mode=2, tgroup=tg)
        np.testing.assert_array_equal(pk["codes"][b, :c].cpu().numpy(), oracle.pack_codes(r["codes"], labels[b][keep], bits, rb))
        ng = (c + tg - 1) // tg
        same_f32(pk["qparams"][b, :ng], r["qparams"])
        want_dense[b, :c] = r["out"]
    y = ops.unpack_dequant_ex(pk, TDT[dt], S, n, sel, sel.labels, bits, opt, tensor=0)
    same_values(y, want_dense, dt)


def test_hidden_state_gather(rt):
    """north star (3): 'coalesced gather of the retained hidden states and KV rows'."""
    _, ops = rt
    rng = np.random.default_rng(3)
    B, S, n = 2, 300, 4096
    sel, labels, ms, rows = selection(ops, rng, B, S, (2, 4, 8))
    for dtype in (torch.float32, torch.bfloat16):
        h = torch.randn(B, S, n, device="cuda").to(dtype)
        g = ops.gather_rows(h, sel, ms)
        mask = sel.mask.bool()
        for b in range(B):
            c = int(rows["n_selected"][b])
            assert torch.equal(g[b, :c], h[b][mask[b]])
            assert int(g[b, c:].abs().sum()) == 0


# ---------------------------------------------------------------- through the compressor
def _layer(B=2, H=8, S=300, HD=512, seed=0):
    from test_parity_gpu import synth_layer
    return synth_layer(B, H, S, HD, seed)


def test_compressor_with_extended_config(rt):
    """key_quant_mode / value_quant_mode / outlier_tau on the config: same selection as the reference mode,
    K/V quantised with the extended kernels (== the oracle on the kept rows)."""
    rtkv, ops = rt
    K, V, A, P = _layer()
    B, S, HD = K.shape
    base = dict(num_hidden_layers=8, num_attention_heads=8, hidden_size=HD, high_precision_bits=8,
                medium_precision_bits=4, low_precision_bits=2)
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    ref = rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig(**base))
    k0, v0, info0 = ref.compress_layer_kv_cache(K, V, A, ids, 6)
    cfg = rtkv.CompressionConfig(**base, key_quant_mode="channel", value_quant_mode="token_group", quant_group_size=64,
                                 quant_token_group=16, outlier_tau=6.0)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    k1, v1, info1 = comp.compress_layer_kv_cache(K, V, A, ids, 6)
    mask = info1["propagation_info"]["selection_mask"].cpu().numpy()
    np.testing.assert_array_equal(mask, info0["propagation_info"]["selection_mask"].cpu().numpy())
    labels = info1["quantization_info"]["bit_assignments"].astype(np.uint8)
    ms = k1.shape[1]
    wk, _ = oracle_dense(K.cpu(), oracle.F32, labels, mask, ms, (2, 4, 8), 2, 0, 16, 0.0)
    wv, _ = oracle_dense(V.cpu(), oracle.F32, labels, mask, ms, (2, 4, 8), 1, 64, 0, 6.0)
    same(k1, wk, oracle.F32)
    same(v1, wv, oracle.F32)
    # finer scales cannot be worse than one scale per row
    kept = torch.from_numpy(mask.astype(bool)).cuda()
    for b in range(B):
        c = int(kept[b].sum())
        e_ref = (v0[b, :c] - V[b][kept[b]]).abs().mean()
        e_ext = (v1[b, :c] - V[b][kept[b]]).abs().mean()
        assert float(e_ext) <= float(e_ref)
    with pytest.raises(NotImplementedError):
        comp.compress_layer_kv_cache(K.cpu(), V.cpu(), A.cpu(), ids.cpu(), 6)


@pytest.mark.parametrize("extended", [False, True])
def test_packed_container_roundtrip(rt, extended):
    rtkv, ops = rt
    K, V, A, P = _layer(seed=3)
    B, S, HD = K.shape
    kw = dict(num_hidden_layers=8, num_attention_heads=8, hidden_size=HD, high_precision_bits=8,
              medium_precision_bits=4, low_precision_bits=2)
    if extended:
        kw.update(key_quant_mode="channel", value_quant_mode="token_group", quant_group_size=128, outlier_tau=5.0)
    comp = rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig(**kw))
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    k, v, info = comp.compress_layer_kv_cache(K, V, A, ids, 2)
    packed = comp.compress_layer_kv_cache_packed(K, V, A, ids, 2, outlier_cap=64)
    dk, dv = packed.dequantize()
    assert packed.max_selected == k.shape[1]
    assert torch.equal(dk, k) and torch.equal(dv, v)          # torch.equal: -0.0 == +0.0
    dense_bytes = 2 * int(info["propagation_info"]["selection_mask"].sum()) * HD * 4
    assert packed.nbytes() < dense_bytes / 3                  # 8/4/2-bit codes vs fp32 rows


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16])
def test_dynamic_range_stress_extended_and_packed(rt, dt):
    """The fast-path edge rows of test_fake_quant_dynamic_range_stress through the group-wise kernel (with and without the
    outlier split) and the packed reference-mode kernel."""
    _, ops = rt
    from test_parity_gpu import stress_rows
    rng = np.random.default_rng(78)
    x = torch.from_numpy(np.stack(stress_rows(rng, 1024))[None]).to(TDT[dt])
    xn, _ = oracle.from_torch(x)
    S, n = x.shape[1], x.shape[2]
    bits = (2, 4, 8)
    labels = rng.integers(0, 3, (1, S)).astype(np.uint8)
    lab_d = torch.from_numpy(labels).cuda()
    for mode, group, tau in (("token_group", 128, 0.0), ("token_group", 128, 6.0), ("token", 0, 6.0)):
        opt = ops.quant_options(mode, mode, group_size=group, outlier_tau=tau)
        y, _ = ops.quant_gather_ex(x.cuda(), None, None, lab_d, bits, S, opt)
        r = oracle.fake_quant_ex(xn[0], dt, labels[0], bits, mode=MODE[mode], group=group, tau=tau)
        same(y, r["out"][None], dt)
    codes, qp = ops.quant_pack(x.cuda(), None, lab_d, bits, S)
    out, ocodes, oqp = oracle.fake_quant_rows(xn, dt, labels, bits, True, True)
    rb = ops.packed_row_bytes(n, max(bits))
    np.testing.assert_array_equal(codes.cpu().numpy().reshape(S, rb), oracle.pack_codes(ocodes, labels, bits, rb))


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
def test_fuzz_extended_and_packed(rt, dt):
    """The random rows of test_fake_quant_fuzz through the group-wise (with / without outliers), per-channel and packed
    kernels, which share the quantiser's row-level fast path."""
    _, ops = rt
    from test_parity_gpu import fuzz_rows
    rng = np.random.default_rng(4242)
    R, n = 600, 512
    x = torch.from_numpy(fuzz_rows(rng, R, n)[None]).to(TDT[dt])
    xn, _ = oracle.from_torch(x)
    bits = (2, 4, 8)
    labels = rng.integers(0, 3, (1, R)).astype(np.uint8)
    lab_d = torch.from_numpy(labels).cuda()
    for mode, group, tg, tau in (("token_group", 128, 0, 0.0), ("token_group", 32, 0, 6.0), ("channel", 0, 16, 0.0),
                                 ("channel", 0, 32, 0.0)):
        opt = ops.quant_options(mode, mode, group_size=group, token_group=tg, outlier_tau=tau)
        y, _ = ops.quant_gather_ex(x.cuda(), None, None, lab_d, bits, R, opt)
        r = oracle.fake_quant_ex(xn[0], dt, labels[0], bits, mode=MODE[mode], group=group, tgroup=tg, tau=tau)
        same(y, r["out"][None], dt)
    rb = ops.packed_row_bytes(n, max(bits))
    codes, qp = ops.quant_pack(x.cuda(), None, lab_d, bits, R)
    out, ocodes, oqp = oracle.fake_quant_rows(xn, dt, labels, bits, True, True)
    np.testing.assert_array_equal(codes.cpu().numpy().reshape(R, rb), oracle.pack_codes(np.nan_to_num(ocodes), labels, bits, rb))
    pk = ops.quant_pack_ex(x.cuda(), None, lab_d, bits, R, ops.quant_options("token_group", group_size=128))
    r = oracle.fake_quant_ex(xn[0], dt, labels[0], bits, mode=1, group=128)
    np.testing.assert_array_equal(pk["codes"][0].cpu().numpy(), oracle.pack_codes(np.nan_to_num(r["codes"]), labels[0], bits, rb))
