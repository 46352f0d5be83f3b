"""GPU parity: every librtkv kernel (through the C ABI) against the CPU oracle on the
same seeded inputs.  Integer / index / code work is compared bit-exactly; the floating
point stages are bit-exact too because oracle and kernels share one defined operation
order (oracle/rtkv_oracle.c header)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import LAYER_CASES, STAGE_CASES, load_golden

pytestmark = pytest.mark.gpu

TDT = {oracle.F32: torch.float32, oracle.F16: torch.float16, oracle.BF16: torch.bfloat16}


@pytest.fixture(scope="module")
def rt():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv
    from rtkv import ops
    name = torch.cuda.get_device_name(0)
    print("device:", name)
    return rtkv, ops


def dev(a, dt=oracle.F32):
    return oracle.to_torch(np.ascontiguousarray(a), dt).cuda()


def host(t):
    if t.dtype in TDT.values():
        return oracle.from_torch(t)[0]
    return t.detach().cpu().numpy()


def assert_same(got, want, dt=oracle.F32):
    """Bit-exact comparison of a device tensor with an oracle array (storage dtype `dt`):
    equal values, equal zero signs, NaN where the oracle has NaN (payloads are not compared)."""
    g = got.detach().float().cpu().numpy()
    w = oracle.to_torch(np.ascontiguousarray(want), dt).float().numpy()
    np.testing.assert_array_equal(g, w)
    ok = ~np.isnan(w)
    np.testing.assert_array_equal(np.signbit(g)[ok], np.signbit(w)[ok])


def causal_probs(B, H, S, gen, dtype=torch.float32, sharp=1.0):
    q = torch.randn(B, H, S, 16, generator=gen)
    k = torch.randn(B, H, S, 16, generator=gen)
    logits = sharp * q @ k.transpose(-1, -2) / 4.0 + torch.full((S, S), float("-inf")).triu(1)
    return torch.softmax(logits, -1, dtype=torch.float32).to(dtype)


# ---------------------------------------------------------------- K1a prompt mass
@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("B,H,S,P", [(1, 8, 64, 12), (2, 32, 700, 128), (1, 40, 333, 66), (1, 4, 17, 3),
                                      (1, 32, 2048, 128), (3, 5, 129, 25), (1, 1, 2, 1)])
def test_prompt_mass_full_matrix(rt, dt, B, H, S, P):
    _, ops = rt
    gen = torch.Generator().manual_seed(S * 31 + H)
    A = causal_probs(B, H, S, gen, TDT[dt])
    An, adt = oracle.from_torch(A)
    want = oracle.prompt_mass(An, adt, P)
    got = ops.prompt_mass(A.cuda(), prompt_len=P)
    np.testing.assert_array_equal(host(got), want)
    # narrow slice gives the same bits (and exercises the other alignment path)
    got2 = ops.prompt_mass(A[..., :P].contiguous().cuda(), prompt_len=P)
    np.testing.assert_array_equal(host(got2), want)


def test_prompt_mass_arbitrary_indices(rt):
    _, ops = rt
    gen = torch.Generator().manual_seed(5)
    A = causal_probs(2, 8, 300, gen)
    idx = torch.tensor([7, 3, 250, 0, 299, 11, 12, 13, 200] + list(range(20, 160)))
    want = oracle.prompt_mass(A.numpy(), oracle.F32, None, idx=idx.numpy())
    got = ops.prompt_mass(A.cuda(), prompt_indices=idx.cuda())
    np.testing.assert_array_equal(host(got), want)


def test_prompt_mass_close_to_torch(rt):
    """Against the reference's own expression (ATen reduction order): a few ulp."""
    _, ops = rt
    gen = torch.Generator().manual_seed(9)
    A = causal_probs(1, 32, 512, gen)
    ref = A[:, :, :, :102].mean(dim=1).sum(dim=-1)
    got = ops.prompt_mass(A.cuda(), prompt_len=102).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-6, atol=1e-7)


# ---------------------------------------------------------------- scores / labels
@pytest.mark.parametrize("adt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("B,S", [(1, 2), (2, 17), (1, 700), (3, 4096), (1, 32768)])
def test_scores_and_labels(rt, k2_mode, adt, B, S):
    rtkv, ops = rt
    rng = np.random.default_rng(S + B)
    mass = rng.random((B, S), dtype=np.float32) * 3
    if adt != oracle.F32:
        mass = oracle.from_torch(torch.from_numpy(mass).to(TDT[adt]).float())[0]
    if B > 1:
        mass[1, :] = 0.25            # degenerate row: max-min <= 1e-8 -> normalised term is 0
    P = max(1, min(S // 5, 128))
    cfg = rtkv.CompressionConfig(num_hidden_layers=32)
    for layer in (0, 17, 31):
        p = ops.make_params(cfg, layer, S, P, 0.8, adt)
        sel = ops.score_select(p, B, S, torch.device("cuda"), mass=torch.from_numpy(mass).cuda(), select=False)
        want = oracle.importance_scores(mass, P, cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer], adt)
        np.testing.assert_array_equal(host(sel.scores), want)
        wl, wc = oracle.assign_precision(want, cfg.theta_h, cfg.theta_m)
        np.testing.assert_array_equal(host(sel.labels), wl)
        _, rows = sel.read_stats()
        assert rows["label_count"].sum(axis=0).tolist() == wc.tolist()
        np.testing.assert_allclose(rows["score_sum"].sum(), want.astype(np.float64).sum(), rtol=1e-12)
        assert rows["score_min"].min() == want.min() and rows["score_max"].max() == want.max()


def test_position_bias_table_one_million(rt):
    """device (float)log((double)x) == host for every x <= 2^20 (the 'correctly rounded logf' contract)."""
    _, ops = rt
    S = 1 << 20
    np.testing.assert_array_equal(host(ops.position_bias(S, torch.device("cuda"))), oracle.position_bias(S))
    for s in (1, 2, 3, 4096, 32768):
        np.testing.assert_array_equal(host(ops.position_bias(s, torch.device("cuda"))), oracle.position_bias(s))


def test_normalize_matches_oracle(rt):
    _, ops = rt
    rng = np.random.default_rng(3)
    m = rng.random((4, 999), dtype=np.float32)
    m[2] = 7.0
    for adt in (oracle.F32, oracle.BF16):
        mm = m if adt == oracle.F32 else oracle.from_torch(torch.from_numpy(m).bfloat16().float())[0]
        np.testing.assert_array_equal(host(ops.normalize_mass(torch.from_numpy(mm).cuda(), adt)), oracle.normalize(mm, adt))


# ---------------------------------------------------------------- K2 selection
@pytest.fixture(params=["single-cta", "cluster"])
def k2_mode(request, rt):
    """Every selection test runs against both K2 kernels (one CTA per row / a cluster of 8 CTAs;
    mode 2 forces the cluster kernel for every row of at least 1024 tokens)."""
    from rtkv import _lib
    _lib.check(_lib.lib.rtkv_debug_k2_mode(1 if request.param == "single-cta" else 2))
    yield request.param
    _lib.check(_lib.lib.rtkv_debug_k2_mode(0))


def run_select(ops, scores, labels, bits, budget, fallback=False, topk=None):
    B, S = scores.shape
    p = ops.LayerParams()
    p.budget = budget
    p.bits[0], p.bits[1], p.bits[2] = bits
    p.fallback_topk = topk if topk is not None else max(1, int(S * 0.1))
    sel = ops.score_select(p, B, S, torch.device("cuda"), scores=torch.from_numpy(scores).cuda(),
                           labels=torch.from_numpy(labels).cuda(), select=True, fallback=fallback)
    layer, rows = sel.read_stats()
    return sel, layer, rows


def check_select(ops, scores, labels, bits, budget):
    sel, layer, rows = run_select(ops, scores, labels, bits, budget)
    mask, cnt, cost, ssum, cc = oracle.select_tokens(scores, labels, bits, budget)
    np.testing.assert_array_equal(host(sel.mask), mask)
    assert rows["n_selected"].tolist() == cnt.tolist()
    assert (rows["cost_bits"] / 8.0).tolist() == cost.tolist()
    assert rows["sel_class"].tolist() == cc.tolist()
    np.testing.assert_allclose(rows["sel_score_sum"], ssum, rtol=1e-10, atol=1e-300)
    idx = host(sel.sel_idx)
    for b in range(scores.shape[0]):
        np.testing.assert_array_equal(idx[b, :cnt[b]], np.nonzero(mask[b])[0])     # ascending original order
    assert int(layer["max_selected"]) == int(cnt.max())
    return cnt


BITS = [(2, 4, 8), (4, 8, 16), (8, 8, 8), (4, 4, 4), (8, 4, 2), (3, 5, 7), (1, 1, 24)]


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("S", [1, 2, 9, 100, 1023, 1024, 1025, 4096, 20000])
def test_select_random_scores(rt, k2_mode, bits, S):
    _, ops = rt
    rng = np.random.default_rng(S * 7 + bits[0])
    B = 3
    scores = rng.random((B, S), dtype=np.float32)
    mono, _ = oracle.assign_precision(scores, 0.7, 0.3)
    rand = rng.integers(0, 3, (B, S)).astype(np.uint8)
    for labels in (mono, rand):
        for ratio in (0.8, 0.6, 0.4, 1.0, 0.05, 3.0):
            check_select(ops, scores, labels, bits, S * ratio)


@pytest.mark.parametrize("S", [40, 1500, 5000, 70000])
def test_select_ties_break_to_lowest_index(rt, k2_mode, S):
    _, ops = rt
    rng = np.random.default_rng(S)
    labels = rng.integers(0, 3, (2, S)).astype(np.uint8)
    # all equal; few distinct values; equal within label classes
    for scores in (np.zeros((2, S), np.float32),
                   rng.integers(0, 4, (2, S)).astype(np.float32) / 4,
                   (labels.astype(np.float32) * 0.4)):
        for bits in ((2, 4, 8), (8, 8, 8)):
            for ratio in (0.8, 0.33):
                check_select(ops, scores, labels, bits, S * ratio)


@pytest.mark.parametrize("S", [3000, 40000])
def test_select_clustered_scores_force_refinement(rt, k2_mode, S):
    """Nearly all scores inside one level-0 bucket: the boundary bucket must be refined."""
    _, ops = rt
    rng = np.random.default_rng(S + 1)
    scores = (0.5 + rng.random((2, S), dtype=np.float32) * 1e-6).astype(np.float32)
    scores[:, 0] = 1e30
    scores[:, 1] = -1e30
    scores[0, 5] = np.inf
    scores[1, 6] = -np.inf
    labels = rng.integers(0, 3, (2, S)).astype(np.uint8)
    for bits in ((2, 4, 8), (4, 4, 4)):
        for ratio in (0.8, 0.4, 0.01):
            check_select(ops, scores, labels, bits, S * ratio)
    # duplicates inside the cluster as well
    scores = np.round(scores * 2 ** 22) / 2 ** 22
    check_select(ops, scores.astype(np.float32), labels, (2, 4, 8), S * 0.6)


def test_select_special_values(rt, k2_mode):
    _, ops = rt
    S = 600
    rng = np.random.default_rng(0)
    scores = rng.standard_normal((1, S)).astype(np.float32)
    scores[0, 10] = np.nan            # torch sorts NaN first
    scores[0, 11] = np.nan
    scores[0, 20] = -0.0
    scores[0, 21] = 0.0               # -0 == +0: tie -> lowest index
    labels = rng.integers(0, 3, (1, S)).astype(np.uint8)
    for budget in (1.0, 5.5, S * 0.3, 0.0, -1.0, float("nan")):
        check_select(ops, scores, labels, (2, 4, 8), budget)


def test_select_budget_is_double(rt, k2_mode):
    _, ops = rt
    S = 4096
    scores = np.linspace(1, 0, S, dtype=np.float32)[None].copy()
    lab = np.ones((1, S), np.uint8)
    lab[0, -10:] = 0
    cnt = check_select(ops, scores, lab, (2, 8, 8), S * 0.8)       # 3276.8 -> 3276 + 3 LOW
    assert cnt[0] == 3279


def test_select_fallback_topk(rt, k2_mode):
    """Nothing fits (cost 2.0 > budget): every row falls back to its top max(1,int(0.1*S)) scores."""
    _, ops = rt
    rng = np.random.default_rng(4)
    for S in (2, 25, 1000):
        scores = rng.random((3, S), dtype=np.float32)
        labels = np.full((3, S), 2, np.uint8)
        sel, layer, rows = run_select(ops, scores, labels, (4, 8, 16), 1.5, fallback=True)
        k = max(1, int(S * 0.1))
        assert int(layer["used_fallback"]) == 1 and int(layer["max_selected"]) == k
        np.testing.assert_array_equal(host(sel.mask), oracle.topk_mask(scores, k))
        assert rows["n_selected"].tolist() == [k] * 3 and rows["n_selected_greedy"].tolist() == [0] * 3
        # one row that does select something disables the fallback for all rows
        labels[1, :] = 0
        sel, layer, rows = run_select(ops, scores, labels, (4, 8, 16), 1.5, fallback=True)
        assert int(layer["used_fallback"]) == 0
        mask, cnt, _, _, _ = oracle.select_tokens(scores, labels, (4, 8, 16), 1.5)
        np.testing.assert_array_equal(host(sel.mask), mask)
        assert rows["n_selected"].tolist() == cnt.tolist() and cnt[0] == 0 and cnt[1] == min(3, S)


def test_cross_row_ticket_is_left_zeroed(rt, k2_mode):
    _, ops = rt
    scores = np.random.default_rng(1).random((5, 300), dtype=np.float32)
    labels = np.ones((5, 300), np.uint8)
    for _ in range(3):
        sel, layer, rows = run_select(ops, scores, labels, (2, 4, 8), 200.0)
        assert layer["reserved"].tolist() == [0, 0]


@pytest.mark.parametrize("H,S,P", [(32, 300, 60), (40, 200, 40), (8, 1000, 128), (16, 77, 15)])
def test_head_sharded_partials_are_bit_identical(rt, H, S, P):
    """ranks holding H/R consecutive heads exchange scalar per-group partials; the combined mass
    equals the unsharded kernel (and the oracle) bit for bit for R = 1, 2, 4, 8."""
    _, ops = rt
    gen = torch.Generator().manual_seed(H * S)
    A = causal_probs(2, H, S, gen)
    want, wparts = oracle.prompt_mass_groups(A.numpy(), oracle.F32, P)
    full = ops.prompt_mass(A.cuda(), prompt_len=P)
    np.testing.assert_array_equal(host(full), want)
    gs = H // 8
    for R in (1, 2, 4, 8):
        hl = H // R
        parts = [ops.prompt_mass_partials(A[:, r * hl:(r + 1) * hl].contiguous().cuda(), P, gs) for r in range(R)]
        gathered = torch.cat(parts, dim=-1)                       # what the all-gather produces
        np.testing.assert_array_equal(host(gathered), wparts)
        np.testing.assert_array_equal(host(ops.mass_from_partials(gathered, H)), want)
        stacked = torch.stack(parts, dim=0)                       # what all_gather_into_tensor produces
        np.testing.assert_array_equal(host(ops.mass_from_partials(stacked, H, rank_major=True)), want)
    np.testing.assert_array_equal(oracle.mass_from_partials(wparts, H), want)


# ---------------------------------------------------------------- K3 quant + gather
@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("n", [8, 64, 96, 257, 1024, 4096, 5120, 8192, 9000])
def test_fake_quant_all_tokens(rt, dt, n):
    _, ops = rt
    rng = np.random.default_rng(n)
    B, S = 2, 37
    x = torch.from_numpy(rng.standard_normal((B, S, n)).astype(np.float32) * rng.choice([0.02, 1.0, 30.0]))
    x[0, 1] = -2.0                       # constant rows: scale=1, zp=0
    x[1, 2] = 0.75
    x[0, 3, 0] = float("inf")
    x[1, 4, 1] = float("nan")
    x[0, 5] = 0.0
    x[1, 6] = -0.3                       # constant row that rounds to -0.0: the reference keeps the sign
    x = x.to(TDT[dt])
    v = (x.float() * 0.5 + 1.0).to(TDT[dt])
    labels = rng.integers(0, 3, (B, S)).astype(np.uint8)
    bits = (2, 4, 8)
    yk, yv = ops.quant_gather(x.cuda(), v.cuda(), None, torch.from_numpy(labels).cuda(), bits, S)
    xn, _ = oracle.from_torch(x)
    vn, _ = oracle.from_torch(v)
    wk, _, _ = oracle.fake_quant_rows(xn, dt, labels, bits)
    wv, _, _ = oracle.fake_quant_rows(vn, dt, labels, bits)
    assert_same(yk, wk, dt)
    assert_same(yv, wv, dt)


def stress_rows(rng, n):
    """fp32 rows sitting on every edge of the quantiser's row-level fast path (see the test below)."""
    rows = []
    for k in range(-45, 46, 3):
        r = rng.standard_normal(n).astype(np.float32) * np.float32(2.0 ** k)
        r[rng.integers(0, n, 40)] = 0.0
        r[rng.integers(0, n, 10)] = -0.0
        r[rng.integers(0, n, 20)] *= np.float32(2.0 ** -40)            # far below the quantisation step
        rows.append(r)
    for k in (0, 20, -20):
        r = np.abs(rng.standard_normal(n).astype(np.float32)) * np.float32(2.0 ** k)
        r[5] = np.float32(-1e-45)                                        # tiny negative minimum: zero-point rounds to 0
        r[6] = np.float32(1e-44)
        rows.append(r)
        rows.append(np.abs(r) + np.float32(2.0 ** k) * 1000)            # all positive, large offset (zp << 0)
        rows.append(-np.abs(r))                                          # all negative (max == -0 possible)
    for k in (57, 59, 60, 61, 63, 90, 120):                              # both sides of the upper window edge 2^60
        r = rng.standard_normal(n).astype(np.float32) * np.float32(2.0 ** k)
        r[rng.integers(0, n, 30)] = 0.0
        rows.append(r)
        rows.append(np.abs(r) * np.float32(2.0 ** -3) + np.float32(2.0 ** k))   # narrow range far from zero
    r = rng.standard_normal(n).astype(np.float32); r[3] = np.inf; rows.append(r)
    r = rng.standard_normal(n).astype(np.float32); r[3] = -np.inf; rows.append(r)
    r = rng.standard_normal(n).astype(np.float32); r[9] = np.nan; rows.append(r)
    rows.append(np.full(n, -0.3, np.float32)); rows.append(np.zeros(n, np.float32))
    rows.append(rng.standard_normal(n).astype(np.float32) * np.float32(1e-42))   # denormal row
    r = rng.standard_normal(n).astype(np.float32) * np.float32(3e38 / 6); rows.append(r)   # range near overflow
    return rows


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
def test_fake_quant_dynamic_range_stress(rt, dt):
    """Rows chosen to sit on every edge of the quantiser's row-level fast path (common.cuh RowQuant): scales from
    2^-45 to 2^45 (inside and outside its window), elements far below the quantisation step, zeros of both signs,
    denormals, a tiny negative minimum whose zero-point rounds to zero, constant, infinite and NaN rows.  Bit-exact
    against the oracle (= the reference's true division and torch.clamp)."""
    _, ops = rt
    rng = np.random.default_rng(77)
    rows = stress_rows(rng, 1024)
    x = torch.from_numpy(np.stack(rows)[None]).to(TDT[dt])
    S = x.shape[1]
    for bits in ((2, 4, 8), (3, 6, 12)):
        labels = rng.integers(0, 3, (1, S)).astype(np.uint8)
        y, _ = ops.quant_gather(x.cuda(), None, None, torch.from_numpy(labels).cuda(), bits, S)
        w, _, _ = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, labels, bits)
        assert_same(y, w, dt)


def fuzz_rows(rng, R, n):
    """R random fp32 rows: magnitudes 2^-40..2^40, ten row shapes (see test_fake_quant_fuzz)."""
    x = np.empty((R, n), np.float32)
    for r in range(R):
        kind = r % 10
        mag = np.float32(2.0 ** rng.uniform(-40, 40))
        if kind == 0: v = rng.standard_normal(n)
        elif kind == 1: v = np.abs(rng.standard_normal(n)) + rng.uniform(0, 5)
        elif kind == 2: v = rng.standard_cauchy(n)
        elif kind == 3: v = rng.standard_normal(n) * (rng.random(n) < 0.1)
        elif kind == 4: v = rng.integers(-50, 50, n).astype(np.float64)
        elif kind == 5: v = np.where(rng.random(n) < 0.5, rng.uniform(-3, 3), rng.uniform(-3, 3))
        elif kind == 6: v = rng.uniform(-3, 3) + rng.standard_normal(n) * 1e-6
        elif kind == 7: v = rng.uniform(100, 1000) * rng.choice([-1, 1]) + rng.random(n)
        elif kind == 8: v = rng.standard_normal(n) * np.where(rng.random(n) < 0.3, 1e-38 / float(mag), 1.0)
        else: v = np.sort(rng.standard_normal(n))
        x[r] = np.clip(v.astype(np.float64) * float(mag), -3e38, 3e38).astype(np.float32)
    return x


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16, oracle.F16])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fake_quant_fuzz(rt, dt, seed):
    """Differential fuzz of the quantiser (fast and plain paths) against the oracle: 1500 rows per case with random
    magnitudes 2^-40..2^40 and ten row shapes (gaussian, one-sided, heavy tailed, sparse, integer valued, two valued, nearly
    constant, tiny range far from zero, denormal mix, sorted ramp), three random bit-width triples per case."""
    _, ops = rt
    rng = np.random.default_rng(1000 + seed)
    R = 1500
    x = fuzz_rows(rng, R, 512)
    xt = torch.from_numpy(x[None]).to(TDT[dt])
    xn = oracle.from_torch(xt)[0]
    for trial in range(3):
        lo, mid, hi = sorted(rng.choice([1, 2, 3, 4, 5, 6, 8, 10, 12] if dt == oracle.F16 else [1, 2, 3, 4, 6, 8, 12, 16], 3))
        labels = rng.integers(0, 3, (1, R)).astype(np.uint8)
        bits = (int(lo), int(mid), int(hi))
        y, _ = ops.quant_gather(xt.cuda(), None, None, torch.from_numpy(labels).cuda(), bits, R)
        w, _, _ = oracle.fake_quant_rows(xn, dt, labels, bits)
        assert_same(y, w, dt)


@pytest.mark.parametrize("dt,bits", [(oracle.F32, (4, 8, 16)), (oracle.BF16, (4, 8, 16)), (oracle.F32, (1, 12, 24)),
                                     (oracle.F16, (3, 5, 11))])
def test_fake_quant_bit_widths(rt, dt, bits):
    _, ops = rt
    rng = np.random.default_rng(bits[2])
    x = torch.from_numpy(rng.standard_normal((1, 64, 512)).astype(np.float32)).to(TDT[dt])
    labels = rng.integers(0, 3, (1, 64)).astype(np.uint8)
    y, _ = ops.quant_gather(x.cuda(), None, None, torch.from_numpy(labels).cuda(), bits, 64)
    w, _, _ = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, labels, bits)
    assert_same(y, w, dt)


@pytest.mark.parametrize("n", [512, 4096, 5120])
def test_bf16_packed_fast_path_equals_elementwise_code(rt, n):
    """bf16 rows: the packed two-elements-per-instruction path (Bf16RowPacked: HADD2/HMUL2.BF16, reciprocal multiply,
    (t+128)-128 rounding) against the element-wise code on the same rows - both are compared with the oracle by the fuzz
    and stress tests in their default setting; this pins them to each other, including rows at the guard's edges."""
    _, ops = rt
    rng = np.random.default_rng(n)
    R = 600
    x = rng.standard_normal((1, R, n)).astype(np.float32)
    x[0, :, :] *= np.exp2(rng.integers(-34, 34, size=(R, 1))).astype(np.float32)       # scales around the guard [2^-30, 2^30]
    x[0, 1] = np.abs(x[0, 1])                      # min == +x
    x[0, 2, 0] = 0.0                                # a zero minimum candidate
    x[0, 3] = np.floor(x[0, 3] * 8) / 8
    x[0, 4, :] = 0.375                              # constant row
    x[0, 5, :n // 2] = 0.0                          # min == 0 exactly
    x[0, 5] = np.abs(x[0, 5])
    xt = torch.from_numpy(x).to(torch.bfloat16).cuda()
    labels = torch.from_numpy(rng.integers(0, 3, (1, R)).astype(np.uint8)).cuda()
    for bits in ((2, 4, 8), (4, 8, 16), (3, 5, 7)):
        ops.lib.rtkv_debug_bf16_packed(1)
        y1, _ = ops.quant_gather(xt, None, None, labels, bits, R)
        ops.lib.rtkv_debug_bf16_packed(0)
        try:
            y0, _ = ops.quant_gather(xt, None, None, labels, bits, R)
        finally:
            ops.lib.rtkv_debug_bf16_packed(1)
        assert torch.equal(y1.view(torch.int16), y0.view(torch.int16)), bits
        want, _, _ = oracle.fake_quant_rows(oracle.from_torch(xt.cpu())[0], oracle.BF16, labels.cpu().numpy(), bits)
        assert_same(y1, want, oracle.BF16)


def test_fp16_sixteen_bits_raises_like_the_reference(rt):
    """torch.clamp(x, 0, 65535) on fp16 raises in the reference (dynamic_quantization.py:121)."""
    rtkv, ops = rt
    x = torch.randn(1, 4, 64, dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError, match="Half"):
        ops.quant_gather(x, None, None, torch.zeros(1, 4, dtype=torch.uint8, device="cuda"), (4, 8, 16), 4)


@pytest.mark.parametrize("dt", [oracle.F32, oracle.BF16])
def test_quant_gather_selected_rows_zero_padded(rt, dt):
    _, ops = rt
    rng = np.random.default_rng(8)
    B, S, n = 3, 200, 256
    scores = rng.random((B, S), dtype=np.float32)
    scores[1] *= 0.2                                   # fewer HIGH tokens -> different counts per row
    labels, _ = oracle.assign_precision(scores, 0.7, 0.3)
    bits = (2, 4, 8)
    sel, layer, rows = run_select(ops, scores, labels, bits, S * 0.25)
    ms = int(layer["max_selected"])
    K = torch.from_numpy(rng.standard_normal((B, S, n)).astype(np.float32)).to(TDT[dt])
    V = torch.from_numpy(rng.standard_normal((B, S, n)).astype(np.float32)).to(TDT[dt])
    yk, yv = ops.quant_gather(K.cuda(), V.cuda(), sel, sel.labels, bits, ms)
    mask = host(sel.mask)
    assert len(set(rows["n_selected"].tolist())) > 1
    for x, y in ((K, yk), (V, yv)):
        q, _, _ = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, labels, bits)
        assert_same(y, oracle.gather_rows(q, dt, mask, ms), dt)
    # plain gather (no quantisation), as selective_propagation.py:229-232 does
    gk, _ = ops.quant_gather(K.cuda(), None, sel, None, None, ms)
    assert_same(gk, oracle.gather_rows(oracle.from_torch(K)[0], dt, mask, ms), dt)
    ss, sl = ops.gather_scalars(sel, ms)
    for b in range(B):
        c = int(rows["n_selected"][b])
        np.testing.assert_array_equal(host(ss)[b, :c], scores[b][mask[b] == 1])
        assert (host(ss)[b, c:] == 0).all() and (sl.cpu().numpy()[b, c:] == 0).all()
        np.testing.assert_array_equal(sl.cpu().numpy()[b, :c], labels[b][mask[b] == 1])


@pytest.mark.parametrize("R", [2, 4, 8])
def test_head_sharded_layer_equals_single_gpu(rt, R):
    """R ranks emulated in one process (collectives replaced by cat / min / max): the concatenated
    rank outputs equal the single-GPU compress_layer_kv_cache bit for bit."""
    rtkv, ops = rt
    H, D, S, B = 16, 32, 400, 2
    K, V, A, P = synth_layer(B, H, S, H * D, seed=R)
    cfg = rtkv.CompressionConfig(num_attention_heads=H, hidden_size=H * D, high_precision_bits=8,
                                 medium_precision_bits=4, low_precision_bits=2)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    layer = 25
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    k_ref, v_ref, info = comp.compress_layer_kv_cache(K, V, A, ids, layer)
    hl, gs, wl = H // R, H // 8, H * D // R
    parts = [ops.prompt_mass_partials(A[:, r * hl:(r + 1) * hl].contiguous(), P, gs) for r in range(R)]
    mass = ops.mass_from_partials(torch.cat(parts, dim=-1).contiguous(), H)
    params, _ = comp.layer_params(layer, S, ops.F32)
    sel = ops.score_select(params, B, S, K.device, mass=mass)
    lyr, rows = sel.read_stats()
    ms = int(lyr["max_selected"])
    assert ms == k_ref.shape[1]
    np.testing.assert_array_equal(host(sel.mask), info["propagation_info"]["selection_mask"].cpu().numpy().astype(np.uint8))
    Ks = [K[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    Vs = [V[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    mms = torch.stack([ops.row_minmax(Ks[r], Vs[r], sel, ms) for r in range(R)])
    mm = torch.stack([mms[..., 0].amin(0), mms[..., 1].amax(0), mms[..., 2].amin(0), mms[..., 3].amax(0)], dim=-1)
    outs = [ops.quant_gather_minmax(Ks[r], Vs[r], sel, sel.labels, (2, 4, 8), ms, mm) for r in range(R)]
    assert torch.equal(torch.cat([o[0] for o in outs], dim=-1), k_ref)
    assert torch.equal(torch.cat([o[1] for o in outs], dim=-1), v_ref)


@pytest.mark.parametrize("kvdt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R,H,D,S,B,bits", [(8, 32, 128, 600, 1, (2, 4, 8)), (4, 32, 128, 257, 2, (4, 8, 16)),
                                            (2, 16, 128, 400, 2, (2, 4, 8)), (8, 16, 24, 333, 2, (2, 4, 8)),
                                            (1, 8, 64, 200, 1, (2, 4, 8)), (4, 8, 64, 9000, 1, (4, 8, 16))])
@pytest.mark.parametrize("variant", [0, 1])
def test_head_sharded_fused_exchange_equals_single_gpu(rt, variant, kvdt, R, H, D, S, B, bits):
    """The one-exchange scheme (rtkv_shard_stats -> rtkv_shard_combine -> K2 -> rtkv_quant_gather_token_minmax), R ranks
    emulated on one device with plain buffers instead of symmetric memory: records, mass, row min/max, selection and
    the concatenated outputs equal the single-GPU path bit for bit; repeated layers exercise the epoch parity."""
    rtkv, ops = rt
    from rtkv import dist as rd
    if kvdt == torch.float16 and max(bits) >= 16:
        pytest.skip("fp16 rows cannot hold 16-bit codes (reference raises)")
    ops.lib.rtkv_debug_shard_variant(variant)        # 0: bulk-copy staged statistics kernel, 1: generic register kernel
    try:
        _fused_exchange_case(rtkv, ops, rd, kvdt, R, H, D, S, B, bits)
    finally:
        ops.lib.rtkv_debug_shard_variant(0)


def _fused_exchange_case(rtkv, ops, rd, kvdt, R, H, D, S, B, bits):
    HD = H * D
    K, V, A, P = synth_layer(B, H, S, HD, seed=R * 7 + S)
    K, V = (K * 3).to(kvdt), (V * 2 + 0.5).to(kvdt)
    K[0, 5] = 1.25                                   # a constant row (scale 1 / zero-point 0 branch)
    cfg = rtkv.CompressionConfig(num_attention_heads=H, hidden_size=HD, high_precision_bits=bits[2],
                                 medium_precision_bits=bits[1], low_precision_bits=bits[0])
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    hl, wl = H // R, HD // R
    Ks = [K[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    Vs = [V[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    As = [A[:, r * hl:(r + 1) * hl].contiguous() for r in range(R)]
    xch = rd.LocalExchange(B, S, R, K.device)
    ranks = [rd.HeadShardedCompressor(comp, H, mode="fused", exchange=xch.view(r)) for r in range(R)]
    want_mass = ops.prompt_mass(A, prompt_len=P)
    want_mm = torch.stack([K.float().amin(-1), K.float().amax(-1), V.float().amin(-1), V.float().amax(-1)], dim=-1)
    for layer in (3, 12, 25, 30, 31):               # odd count: both parities, and a wrap
        k_ref, v_ref, info = comp.compress_layer_kv_cache(K, V, A, ids, layer)
        for r in range(R):
            ranks[r].fused_publish(Ks[r], Vs[r], As[r])
        outs = []
        for r in range(R):
            sel, mm = ranks[r].fused_select(Ks[r], layer)
            assert torch.equal(ranks[r].exchange.mass, want_mass) and torch.equal(mm, want_mm)
            lyr, rows = sel.read_stats()
            ms = int(lyr["max_selected"])
            assert ms == k_ref.shape[1]
            assert torch.equal(sel.mask.view(torch.bool), info["propagation_info"]["selection_mask"])
            outs.append(ranks[r]._fused_quantise(Ks[r], Vs[r], sel, mm, ms, True))
            assert int(xch.states[r][2]) == 0
        assert torch.equal(torch.cat([o[0] for o in outs], dim=-1), k_ref)
        assert torch.equal(torch.cat([o[1] for o in outs], dim=-1), v_ref)
    # fixed-capacity variant (what bench.py and CUDA graphs use): rows [0, n_selected) equal, nothing else written
    for r in range(R):
        ranks[r].fused_publish(Ks[r], Vs[r], As[r])
    for r in range(R):
        ok, ov = torch.full((B, S, wl), 7.0, device="cuda", dtype=kvdt), torch.full((B, S, wl), 7.0, device="cuda", dtype=kvdt)
        sel, mm = ranks[r].fused_select(Ks[r], 31)
        ranks[r]._fused_quantise(Ks[r], Vs[r], sel, mm, S, False, ok, ov)
        _, rows = sel.read_stats()
        for b in range(B):
            c = int(rows["n_selected"][b])
            assert torch.equal(ok[b, :c], k_ref[b, :c, r * wl:(r + 1) * wl]) and torch.equal(ov[b, :c], v_ref[b, :c, r * wl:(r + 1) * wl])
            assert bool((ok[b, c:] == 7.0).all()) and bool((ov[b, c:] == 7.0).all())


@pytest.mark.parametrize("kvdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,H,D,S,B", [(8, 32, 128, 2100, 1), (4, 32, 128, 777, 2), (2, 16, 128, 1500, 1)])
def test_head_sharded_one_kernel_exchange_all_ranks_cooperative(rt, kvdt, R, H, D, S, B):
    """rtkv_shard_stats_combine (statistics + peer stores + per-CTA flags + reduction in ONE kernel, the default of real
    multi-GPU runs) with all R ranks inside ONE cooperative launch on this GPU (blockIdx.y = rank): mass and whole-row
    min/max equal the single-GPU values on every rank, over several layers (epoch parity)."""
    rtkv, ops = rt
    from rtkv import dist as rd
    HD = H * D
    K, V, A, P = synth_layer(B, H, S, HD, seed=R + S)
    K, V = K.to(kvdt), V.to(kvdt)
    hl, wl = H // R, HD // R
    Ks = [K[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    Vs = [V[..., r * wl:(r + 1) * wl].contiguous() for r in range(R)]
    As = [A[:, r * hl:(r + 1) * hl].contiguous() for r in range(R)]
    xch = rd.LocalExchange(B, S, R, K.device)
    views = [xch.view(r) for r in range(R)]
    want_mass = ops.prompt_mass(A, prompt_len=P)
    want_mm = torch.stack([K.float().amin(-1), K.float().amax(-1), V.float().amin(-1), V.float().amax(-1)], dim=-1)
    ranges = [ops.new_mass_range(B, K.device) for _ in range(R)]
    cfg = rtkv.CompressionConfig(num_attention_heads=H, hidden_size=HD, high_precision_bits=16, medium_precision_bits=8,
                                 low_precision_bits=4)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    for layer in range(5):
        for v in views:
            v.mass.fill_(-1.0); v.minmax.fill_(-1.0)
        ops.shard_stats_combine_allranks(As, Ks, Vs, P, H // 8, xch.ptrs, [v.state for v in views], H,
                                         [v.mass for v in views], [v.minmax for v in views], ranges)
        torch.cuda.synchronize()
        params, _ = comp.layer_params(layer * 6, S, ops.F32)
        want_by_mode = {}
        for mode in (1, 2):                                 # 1: single-CTA selection kernel, 2: cluster kernel
            ops.lib.rtkv_debug_k2_mode(mode)
            want_by_mode[mode] = ops.score_select(params, B, S, K.device, mass=want_mass)
        ops.lib.rtkv_debug_k2_mode(0)
        for r in range(R):
            assert int(views[r].state[2]) == 0, "a CTA gave up waiting for its peers"
            assert torch.equal(views[r].mass, want_mass) and torch.equal(views[r].minmax, want_mm), (layer, r)
            # the mass range reduced by the exchange kernel lets K2 skip its own pass: same selection, range reset afterwards
            mode = 1 + (layer + r) % 2                             # alternate the two selection kernels
            want_sel = want_by_mode[mode]
            ops.lib.rtkv_debug_k2_mode(mode)
            try:
                sel = ops.score_select(params, B, S, K.device, mass=views[r].mass, mass_range=ranges[r])
            finally:
                ops.lib.rtkv_debug_k2_mode(0)
            assert torch.equal(sel.scores, want_sel.scores) and torch.equal(sel.labels, want_sel.labels)
            assert torch.equal(sel.mask, want_sel.mask) and torch.equal(sel.stats, want_sel.stats)
            assert ranges[r].cpu().tolist() == [[-1, 0, 0, 0]] * B


def test_head_sharded_fused_exchange_propagates_nan_rows(rt):
    """torch.min/max propagate NaN: a NaN anywhere in a row (on any rank) makes that row's scale NaN on every rank."""
    rtkv, ops = rt
    from rtkv import dist as rd
    R, H, D, S, B = 4, 8, 32, 130, 1
    K, V, A, P = synth_layer(B, H, S, H * D, seed=99)
    K[0, 7, 3] = float("nan")            # rank 0's slice
    V[0, 9, H * D - 1] = float("nan")    # rank 3's slice
    cfg = rtkv.CompressionConfig(num_attention_heads=H, hidden_size=H * D, high_precision_bits=8, medium_precision_bits=4,
                                 low_precision_bits=2)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    k_ref, v_ref, _ = comp.compress_layer_kv_cache(K, V, A, torch.zeros(B, S, dtype=torch.long, device="cuda"), 2)
    wl, hl = H * D // R, H // R
    xch = rd.LocalExchange(B, S, R, K.device)
    ranks = [rd.HeadShardedCompressor(comp, H, mode="fused", exchange=xch.view(r)) for r in range(R)]
    sl = [(K[..., r * wl:(r + 1) * wl].contiguous(), V[..., r * wl:(r + 1) * wl].contiguous(), A[:, r * hl:(r + 1) * hl].contiguous())
          for r in range(R)]
    for r in range(R):
        ranks[r].fused_publish(*sl[r])
    ks, vs = [], []
    for r in range(R):
        sel, mm = ranks[r].fused_select(sl[r][0], 2)
        lyr, _ = sel.read_stats()
        k, v = ranks[r]._fused_quantise(sl[r][0], sl[r][1], sel, mm, int(lyr["max_selected"]), True)
        ks.append(k); vs.append(v)
    got_k, got_v = torch.cat(ks, -1), torch.cat(vs, -1)
    assert torch.equal(torch.isnan(got_k), torch.isnan(k_ref)) and torch.equal(torch.isnan(got_v), torch.isnan(v_ref))
    assert bool(torch.isnan(k_ref[0]).any()) and torch.equal(torch.nan_to_num(got_k), torch.nan_to_num(k_ref))
    assert torch.equal(torch.nan_to_num(got_v), torch.nan_to_num(v_ref))


def test_whole_tensor_params_and_quantize(rt):
    rtkv, ops = rt
    q = rtkv.DynamicPrecisionQuantizer(rtkv.CompressionConfig())
    rng = np.random.default_rng(2)
    for dt in (oracle.F32, oracle.BF16):
        x = torch.from_numpy(rng.standard_normal((1, 1, 3000)).astype(np.float32)).to(TDT[dt])
        for bits in (4, 8):
            scale, zp = q.get_quantization_params(x.cuda(), bits)
            out, codes, qp = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, np.zeros((1, 1), np.uint8),
                                                    (bits, bits, bits), True, True)
            assert float(scale) == qp[0, 0] and float(zp) == qp[0, 1]
            y = q.quantize_tensor(x.cuda(), bits, scale, zp)
            assert_same(y, out, dt)
    const = torch.full((5, 7), 3.0, device="cuda")
    s, z = q.get_quantization_params(const, 8)
    assert float(s) == 1.0 and float(z) == 0.0


# ---------------------------------------------------------------- packed rows
def test_packed_bf16_sixteen_bits_unsupported(rt):
    """bf16 rounds the clamp bound 65535 to 65536, which has no 16-bit code."""
    _, ops = rt
    x = torch.randn(1, 4, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="bf16"):
        ops.quant_pack(x, None, torch.zeros(1, 4, dtype=torch.uint8, device="cuda"), (4, 8, 16), 4)


@pytest.mark.parametrize("dt,bits", [(oracle.F32, (2, 4, 8)), (oracle.F32, (4, 8, 16)), (oracle.F32, (1, 2, 4)),
                                     (oracle.BF16, (2, 4, 8)), (oracle.BF16, (4, 8, 12)), (oracle.F16, (1, 8, 15))])
def test_packed_codes_bit_exact_and_roundtrip(rt, dt, bits):
    _, ops = rt
    rng = np.random.default_rng(bits[0] + 100)
    B, S, n = 2, 50, 200
    x = torch.from_numpy(rng.standard_normal((B, S, n)).astype(np.float32)).to(TDT[dt])
    labels = rng.integers(0, 3, (B, S)).astype(np.uint8)
    lab_d = torch.from_numpy(labels).cuda()
    codes, qp = ops.quant_pack(x.cuda(), None, lab_d, bits, S)
    out, ocodes, oqp = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, labels, bits, True, True)
    rb = ops.packed_row_bytes(n, max(bits))
    want = oracle.pack_codes(ocodes, labels, bits, rb)
    np.testing.assert_array_equal(codes.cpu().numpy().reshape(B * S, rb), want)     # integer codes: bit-exact
    np.testing.assert_array_equal(qp.cpu().numpy().reshape(B * S, 2), oqp)
    y = ops.unpack_dequant(codes, qp, TDT[dt], S, n, None, lab_d, bits)
    assert_same(y, out, dt)                                                         # dequantised == fake-quant path


# ---------------------------------------------------------------- whole layer, golden fixtures
def build_compressor(rtkv, m):
    cfg = rtkv.CompressionConfig(
        num_hidden_layers=m["L"], num_attention_heads=m["H"], hidden_size=m["HD"],
        alpha=m["alpha"], beta=m["beta"], gamma=m["gamma"], theta_h=m["theta_h"], theta_m=m["theta_m"],
        high_precision_bits=m["bits"][2], medium_precision_bits=m["bits"][1], low_precision_bits=m["bits"][0],
        early_layer_ratio=m["early"], middle_layer_ratio=m["middle"], later_layer_ratio=m["later"])
    return rtkv.RealTimePrefillCompressor(cfg)


def ulp(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("on_host", [False, True], ids=["cuda-tensors", "host-tensors"])
@pytest.mark.parametrize("name", LAYER_CASES)
def test_golden_layer_through_drop_in_api(rt, name, on_host):
    """The fixture outputs were produced by the unmodified reference."""
    rtkv, ops = rt
    g = load_golden(name)
    m = g["meta"]
    comp = build_compressor(rtkv, m)
    place = (lambda t: t) if on_host else (lambda t: t.cuda())
    K, V, A = (place(oracle.to_torch(g[k], m["dt"])) for k in ("K", "V", "A"))
    ids = place(torch.zeros(m["B"], m["S"], dtype=torch.long))
    k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, m["layer_idx"])
    assert k_sel.device == K.device and k_sel.dtype == K.dtype
    assert list(info["compressed_shape"]) == m["compressed_shape"] and list(k_sel.shape) == m["compressed_shape"]
    np.testing.assert_array_equal(info["propagation_info"]["selection_mask"].cpu().numpy().astype(np.uint8), g["mask"])
    np.testing.assert_array_equal(info["quantization_info"]["bit_assignments"], g["labels"].astype(np.int64))
    assert_same(k_sel, g["k_sel"], m["dt"])
    assert_same(v_sel, g["v_sel"], m["dt"])
    sc = comp.importance_tracker.layer_scores[m["layer_idx"]].numpy()
    assert ulp(sc, g["scores"]).max() <= 8
    assert info["precision_stats"] == m["precision_stats"]
    assert abs(info["compression_ratio"] - m["compression_ratio"]) < 1e-12
    for k in ("mean_score", "std_score", "min_score", "max_score"):
        a, b = info["importance_stats"][k], m["importance_stats"][k]
        assert (a != a and b != b) or abs(a - b) <= 2e-6 * max(1.0, abs(b)), k
    st = info["propagation_info"]["selection_stats"]
    assert st["selected_counts"] == m["selected_counts"] and st["cost_distribution"] == m["cost_distribution"]
    np.testing.assert_allclose(st["budget_utilization"], m["budget_utilization"], rtol=1e-12)
    np.testing.assert_allclose(st["avg_importance"], m["avg_importance"], rtol=1e-5)
    assert info["propagation_info"]["max_selected_length"] == m["compressed_shape"][1]
    assert set(info) == {"layer_idx", "processing_time", "original_shape", "compressed_shape", "compression_ratio",
                         "memory_savings", "importance_stats", "precision_stats", "quantization_info",
                         "propagation_info"}
    overall = comp.get_overall_compression_stats()
    assert overall["total_layers_processed"] == 1
    assert overall["cumulative_compression"] == m["compressed_shape"][1] / m["S"]


@pytest.mark.parametrize("how", ["cuda-tensors", "host-tensors", "async"])
def test_fp16_default_widths_raise_only_when_the_wide_class_is_populated(rt, how):
    """fp16 K/V with the reference's default 16/8/4: the reference raises (torch.clamp overflow, dynamic_quantization.py:121)
    only if a HIGH token exists (:165 skips empty classes).  Fixture made by the reference (oracle/pin_fp16_wide.py); the
    populated-class counterpart of the layer_f16_16_8_4_nohigh golden layer."""
    rtkv, ops = rt
    g = load_golden("layer_f16_16_8_4_high_raises")
    m = g["meta"]
    cfg = rtkv.CompressionConfig(num_hidden_layers=m["L"], num_attention_heads=m["H"], hidden_size=m["HD"])   # 16/8/4
    comp = rtkv.RealTimePrefillCompressor(cfg)
    place = (lambda t: t) if how == "host-tensors" else (lambda t: t.cuda())
    K, V, A = (place(oracle.to_torch(g[k], m["dt"])) for k in ("K", "V", "A"))
    ids = place(torch.zeros(m["B"], m["S"], dtype=torch.long))
    if how == "async":
        ok, ov = torch.empty_like(K), torch.empty_like(V)
        sel = comp.compress_layer_kv_cache_async(K, V, A, m["layer_idx"], ok, ov)       # nothing is read back here ...
        with pytest.raises(RuntimeError, match="Half"):
            comp.collect_async([(m["layer_idx"], sel, ok, ov)])                          # ... the check is deferred to here
        # and a layer without HIGH tokens goes through the same async path cleanly
        g2 = load_golden("layer_f16_16_8_4_nohigh")
        m2 = g2["meta"]
        comp2 = build_compressor(rtkv, m2)
        K2, V2, A2 = (oracle.to_torch(g2[k], m2["dt"]).cuda() for k in ("K", "V", "A"))
        ok2, ov2 = torch.empty_like(K2), torch.empty_like(V2)
        sel2 = comp2.compress_layer_kv_cache_async(K2, V2, A2, m2["layer_idx"], ok2, ov2)
        res = comp2.collect_async([(m2["layer_idx"], sel2, ok2, ov2)])
        k_sel, v_sel, _ = res[m2["layer_idx"]]
        assert_same(k_sel, g2["k_sel"], m2["dt"])
        assert_same(v_sel, g2["v_sel"], m2["dt"])
        return
    with pytest.raises(RuntimeError, match="Half"):
        comp.compress_layer_kv_cache(K, V, A, ids, m["layer_idx"])


@pytest.mark.parametrize("name", STAGE_CASES)
def test_golden_stages_through_component_api(rt, name):
    rtkv, ops = rt
    g = load_golden(name)
    m = g["meta"]
    cfg = rtkv.CompressionConfig(num_hidden_layers=m["L"], high_precision_bits=m["bits"][2],
                                 medium_precision_bits=m["bits"][1], low_precision_bits=m["bits"][0])
    quant, prop = rtkv.DynamicPrecisionQuantizer(cfg), rtkv.SelectiveTokenPropagator(cfg)
    scores = torch.from_numpy(g["scores"])
    K, V = oracle.to_torch(g["K"], m["dt"]), oracle.to_torch(g["V"], m["dt"])
    labels, stats = quant.assign_precision_levels(scores)
    assert labels.dtype == torch.long and labels.device == scores.device and stats == m["precision_stats"]
    np.testing.assert_array_equal(labels.numpy().astype(np.uint8), g["labels"])
    kq, vq, qi = quant.apply_mixed_precision_quantization(K, V, labels)
    assert_same(kq, g["k_quant"], m["dt"])
    assert_same(vq, g["v_quant"], m["dt"])
    assert qi["scales"] == {} and qi["zero_points"] == {} and "bit_assignments" in qi
    mm, _ = prop.select_tokens_with_budget(scores, labels, m["ratio"], m["ratio_layer"])
    np.testing.assert_array_equal(mm.numpy().astype(np.uint8), g["mask_mono"])
    rl = torch.from_numpy(g["rand_labels"].astype(np.int64))
    mr, _ = prop.select_tokens_with_budget(scores, rl, m["ratio"], m["ratio_layer"])
    assert mr.dtype == torch.bool
    np.testing.assert_array_equal(mr.numpy().astype(np.uint8), g["mask_rand"])
    ks, vs, ss, ls, pinfo = prop.apply_token_selection(K, V, scores, rl, m["ratio_layer"])
    assert_same(ks, g["k_sel"], m["dt"])
    assert_same(vs, g["v_sel"], m["dt"])
    np.testing.assert_array_equal(ss.numpy(), g["sel_scores"])
    np.testing.assert_array_equal(ls.numpy().astype(np.uint8), g["sel_labels"])
    assert ls.dtype == torch.long and pinfo["max_selected_length"] == g["k_sel"].shape[1]


# ---------------------------------------------------------------- BASELINE.json sizes
def synth_layer(B, H, S, HD, seed, dtype=torch.float32):
    """BASELINE configs[0] recipe: K,V,Q ~ N(0,1), A = causal softmax(QK^T/sqrt(d)); only the prompt
    columns of A are materialised."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = HD // H
    K = torch.randn(B, H, S, D, generator=g, device="cuda")
    V = torch.randn(B, H, S, D, generator=g, device="cuda")
    Q = torch.randn(B, H, S, D, generator=g, device="cuda")
    P = max(1, min(S // 5, 128))
    logits = Q @ K.transpose(-1, -2) / D ** 0.5
    logits = logits + torch.full((S, S), float("-inf"), device="cuda").triu(1)
    A = torch.softmax(logits, -1)[..., :P].contiguous().to(dtype)
    Kf = K.transpose(1, 2).reshape(B, S, HD).contiguous().to(dtype)
    Vf = V.transpose(1, 2).reshape(B, S, HD).contiguous().to(dtype)
    return Kf, Vf, A, P


@pytest.mark.parametrize("bits", [(2, 4, 8), (4, 8, 16)])
def test_config1_single_layer_full_size_vs_oracle(rt, bits):
    """configs[0]: K/V [1,32,2048,128] fp32 -> [1,2048,4096], P=128, layers 0/15/31: every output bit-exact."""
    rtkv, ops = rt
    K, V, A, P = synth_layer(1, 32, 2048, 4096, seed=0)
    cfg = rtkv.CompressionConfig(high_precision_bits=bits[2], medium_precision_bits=bits[1], low_precision_bits=bits[0])
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(1, 2048, dtype=torch.long, device="cuda")
    Kn, Vn, An = K.cpu().numpy(), V.cpu().numpy(), A.cpu().numpy()
    for layer in (0, 15, 31):
        k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, layer)
        o = oracle.compress_layer(Kn, Vn, oracle.F32, An, oracle.F32, P, cfg.alpha, cfg.beta, cfg.gamma,
                                  cfg.layer_weights[layer], cfg.theta_h, cfg.theta_m, bits,
                                  comp.propagator.get_layer_propagation_ratio(layer))
        np.testing.assert_array_equal(comp.importance_tracker.layer_scores[layer].numpy(), o["scores"])
        np.testing.assert_array_equal(info["propagation_info"]["selection_mask"].cpu().numpy().astype(np.uint8), o["mask"])
        np.testing.assert_array_equal(k_sel.cpu().numpy(), o["k_sel"])
        np.testing.assert_array_equal(v_sel.cpu().numpy(), o["v_sel"])


@pytest.mark.parametrize("seed", range(6))
def test_whole_layer_fuzz(rt, seed):
    """Random layers through the drop-in call against oracle.compress_layer: random batch / heads / length / width,
    dtype, attention sharpness, weights alpha-gamma, thresholds, bit widths and depth band; scores, labels, mask and the
    compressed K/V are compared bit for bit (5 layers per seed)."""
    rtkv, ops = rt
    rng = np.random.default_rng(500 + seed)
    for case in range(5):
        B = int(rng.integers(1, 4)); H = int(rng.choice([1, 3, 8, 16, 40])); S = int(rng.choice([1, 2, 7, 64, 333, 1500, 2600]))
        HD = int(rng.choice([8, 96, 256, 1024, 5120])); L = int(rng.choice([2, 5, 32]))
        dt = int(rng.choice([oracle.F32, oracle.BF16, oracle.F16]))
        bits = tuple(int(b) for b in sorted(rng.choice([1, 2, 3, 4, 6, 8, 12] if dt == oracle.F16 else [1, 2, 4, 8, 12, 16], 3)))
        cfg = rtkv.CompressionConfig(
            num_hidden_layers=L, num_attention_heads=H, hidden_size=HD, alpha=float(rng.uniform(0, 1)),
            beta=float(rng.uniform(0, 1)), gamma=float(rng.uniform(0, 1)), theta_h=float(rng.uniform(0.4, 0.9)),
            theta_m=float(rng.uniform(0.0, 0.4)), early_layer_ratio=float(rng.uniform(0.3, 1.0)),
            middle_layer_ratio=float(rng.uniform(0.2, 0.9)), later_layer_ratio=float(rng.uniform(0.05, 0.6)),
            low_precision_bits=bits[0], medium_precision_bits=bits[1], high_precision_bits=bits[2])
        comp = rtkv.RealTimePrefillCompressor(cfg)
        layer = int(rng.integers(0, L))
        g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
        P = comp.prompt_length(S)
        sharp = float(rng.choice([0.2, 1.0, 6.0]))
        A = torch.softmax(torch.randn(B, H, S, S, generator=g, device="cuda") * sharp +
                          torch.triu(torch.full((S, S), -1e9, device="cuda"), 1), dim=-1).to(TDT[dt])
        K = (torch.randn(B, S, HD, generator=g, device="cuda") * float(2.0 ** rng.uniform(-6, 6))).to(TDT[dt])
        V = (torch.randn(B, S, HD, generator=g, device="cuda") * float(2.0 ** rng.uniform(-6, 6)) + float(rng.uniform(-2, 2))).to(TDT[dt])
        ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
        k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, layer)
        o = oracle.compress_layer(host(K), host(V), dt, host(A), dt, P, cfg.alpha, cfg.beta, cfg.gamma,
                                  cfg.layer_weights[layer], cfg.theta_h, cfg.theta_m, bits,
                                  comp.propagator.get_layer_propagation_ratio(layer))
        tag = f"seed {seed} case {case}: B={B} H={H} S={S} HD={HD} dt={dt} bits={bits} layer={layer}/{L}"
        np.testing.assert_array_equal(comp.importance_tracker.layer_scores[layer].numpy(), o["scores"], err_msg=tag)
        np.testing.assert_array_equal(info["quantization_info"]["bit_assignments"].astype(np.uint8), o["labels"], err_msg=tag)
        np.testing.assert_array_equal(info["propagation_info"]["selection_mask"].cpu().numpy().astype(np.uint8), o["mask"], err_msg=tag)
        assert k_sel.shape[1] == o["max_sel"], tag
        assert_same(k_sel, o["k_sel"], dt)
        assert_same(v_sel, o["v_sel"], dt)


@pytest.mark.parametrize("seed", range(4))
def test_host_path_fuzz(rt, seed):
    """rtkv_compress_layer_host (pieces of K/V uploaded, quantised and downloaded as they arrive; strided attention upload)
    against the CUDA-tensor path on random shapes: identical tensors and statistics."""
    rtkv, ops = rt
    rng = np.random.default_rng(900 + seed)
    for case in range(6):
        B = int(rng.integers(1, 6)); H = int(rng.choice([1, 4, 8])); S = int(rng.choice([1, 5, 63, 500, 2000]))
        HD = int(rng.choice([8, 100, 512, 1024])); L = 8
        dtype = [torch.float32, torch.bfloat16, torch.float16][int(rng.integers(0, 3))]
        cfg = rtkv.CompressionConfig(num_hidden_layers=L, num_attention_heads=H, hidden_size=HD, high_precision_bits=8,
                                     medium_precision_bits=4, low_precision_bits=2,
                                     early_layer_ratio=float(rng.uniform(0.1, 1.0)), theta_h=float(rng.uniform(0.4, 0.9)),
                                     theta_m=float(rng.uniform(0.0, 0.4)))
        dev_c, host_c = rtkv.RealTimePrefillCompressor(cfg), rtkv.RealTimePrefillCompressor(cfg)
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        P = dev_c.prompt_length(S)
        W = int(rng.choice([P, S]))                         # a [..., :P] slice or the full matrix (strided upload)
        A = torch.softmax(torch.randn(B, H, S, S, generator=g) * 3.0 + torch.triu(torch.full((S, S), -1e9), 1), dim=-1)
        for b in range(B):
            A[b] = A[b] * float(rng.uniform(0.2, 1.0))      # different counts per batch row
        A = A.to(dtype)[..., :W]
        K = torch.randn(B, S, HD, generator=g).to(dtype)
        V = torch.randn(B, S, HD, generator=g).to(dtype)
        ids = torch.zeros(B, S, dtype=torch.long)
        kd, vd, idv = dev_c.compress_layer_kv_cache(K.cuda(), V.cuda(), A.cuda(), ids.cuda(), 1)
        kh, vh, ih = host_c.compress_layer_kv_cache(K, V, A, ids, 1)
        tag = f"seed {seed} case {case}: B={B} H={H} S={S} HD={HD} {dtype} W={W}"
        assert not kh.is_cuda and kh.shape == kd.shape, tag
        assert torch.equal(kh.view(torch.uint8), kd.cpu().view(torch.uint8)), tag
        assert torch.equal(vh.view(torch.uint8), vd.cpu().view(torch.uint8)), tag
        assert torch.equal(ih["propagation_info"]["selection_mask"].cpu(), idv["propagation_info"]["selection_mask"].cpu()), tag
        for key in ("compression_ratio", "importance_stats", "precision_stats"):
            assert repr(ih[key]) == repr(idv[key]), (tag, key)          # repr: nan == nan (std of a single token)
        assert repr(ih["propagation_info"]["selection_stats"]) == repr(idv["propagation_info"]["selection_stats"]), tag


def test_config2_layer_4k_uniform_scores_all_widths(rt):
    """configs[1] shape, stage-level run with scores ~ U[0,1) so that all three widths occur
    (SURVEY.md §8d), checked against the oracle on the full [1,4096,4096] tensors."""
    rtkv, ops = rt
    rng = np.random.default_rng(42)
    S, HD, bits = 4096, 4096, (2, 4, 8)
    scores = rng.random((1, S), dtype=np.float32)
    K = torch.randn(1, S, HD, device="cuda")
    V = torch.randn(1, S, HD, device="cuda") * 3 - 1
    p = ops.LayerParams()
    p.theta_h, p.theta_m, p.budget = 0.7, 0.3, S * 0.6
    p.bits[0], p.bits[1], p.bits[2] = bits
    sel = ops.score_select(p, 1, S, K.device, scores=torch.from_numpy(scores).cuda())
    layer, rows = sel.read_stats()
    ms = int(layer["max_selected"])
    yk, yv = ops.quant_gather(K, V, sel, sel.labels, bits, ms)
    labels, _ = oracle.assign_precision(scores, 0.7, 0.3)
    mask, cnt, _, _, _ = oracle.select_tokens(scores, labels, bits, S * 0.6)
    np.testing.assert_array_equal(host(sel.mask), mask)
    assert ms == cnt[0] and len(set(labels[0][mask[0] == 1].tolist())) == 3
    for x, y in ((K, yk), (V, yv)):
        q, _, _ = oracle.fake_quant_rows(x.cpu().numpy(), oracle.F32, labels, bits)
        np.testing.assert_array_equal(y.cpu().numpy(), oracle.gather_rows(q, oracle.F32, mask, ms))


@pytest.mark.parametrize("S,HD,H", [(32768, 4096, 32), (16384, 5120, 40)])
def test_long_context_properties(rt, S, HD, H):
    """configs[2]/[3] sizes: size-independent properties of the device result."""
    rtkv, ops = rt
    g = torch.Generator(device="cuda").manual_seed(S)
    P = 128
    A = torch.rand(1, H, S, P, generator=g, device="cuda") / P
    K = torch.randn(1, S, HD, generator=g, device="cuda")
    V = torch.randn(1, S, HD, generator=g, device="cuda")
    bits = (2, 4, 8)
    cfg = rtkv.CompressionConfig(num_hidden_layers=H, num_attention_heads=H, hidden_size=HD,
                                 high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    layer = H - 1
    k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, torch.zeros(1, S, dtype=torch.long, device="cuda"), layer)
    mask = info["propagation_info"]["selection_mask"][0]
    n = int(mask.sum())
    assert k_sel.shape == (1, n, HD) and info["propagation_info"]["selection_stats"]["selected_counts"] == [n]
    scores = comp.importance_tracker.layer_scores[layer]
    labels = torch.from_numpy(info["quantization_info"]["bit_assignments"])
    # selection == oracle walk (the walk is O(S log S) on the host: fine at this size)
    om, ocnt, ocost, _, _ = oracle.select_tokens(scores.numpy(), labels.numpy().astype(np.uint8), bits, S * 0.4)
    np.testing.assert_array_equal(mask.cpu().numpy().astype(np.uint8), om[0])
    assert info["propagation_info"]["selection_stats"]["budget_utilization"][0] <= 1.0
    # kept rows: each output row is the fake-quantised input row; codes lie on the row's own grid
    idx = torch.nonzero(mask).flatten()
    take = idx[:: max(1, n // 64)]
    pos = torch.arange(n, device="cuda")[:: max(1, n // 64)]
    q, _, _ = oracle.fake_quant_rows(K[0, take].cpu().numpy()[None], oracle.F32,
                                     labels[0, take.cpu()].numpy().astype(np.uint8)[None], bits)
    np.testing.assert_array_equal(k_sel[0, pos].cpu().numpy(), q[0])
    # idempotence of the row gather: compressing with ratio 1.0 keeps every token in order
    cfg1 = rtkv.CompressionConfig(num_hidden_layers=H, num_attention_heads=H, hidden_size=HD, early_layer_ratio=1.0,
                                  middle_layer_ratio=1.0, later_layer_ratio=1.0, high_precision_bits=8,
                                  medium_precision_bits=8, low_precision_bits=8)
    c1 = rtkv.RealTimePrefillCompressor(cfg1)
    k1, _, i1 = c1.compress_layer_kv_cache(K, V, A, torch.zeros(1, S, dtype=torch.long, device="cuda"), 0)
    assert k1.shape == K.shape and bool(i1["propagation_info"]["selection_mask"].all())
    err = (k1 - K).abs().amax(dim=-1)
    rng_ = K.amax(dim=-1) - K.amin(dim=-1)
    assert bool((err <= rng_ / 255 * 0.5001 + 1e-6).all())                 # within half a quantisation step


def test_async_variant_matches_sync(rt):
    rtkv, ops = rt
    K, V, A, P = synth_layer(2, 8, 640, 512, seed=3)
    cfg = rtkv.CompressionConfig(high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(2, 640, dtype=torch.long, device="cuda")
    k_sel, v_sel, info = comp.compress_layer_kv_cache(K, V, A, ids, 12)
    ok, ov = torch.full_like(K, 7.0), torch.full_like(V, 7.0)
    sel = comp.compress_layer_kv_cache_async(K, V, A, 12, ok, ov, zero_pad=True)
    layer, rows = sel.read_stats()
    ms = int(layer["max_selected"])
    assert torch.equal(ok[:, :ms], k_sel) and torch.equal(ov[:, :ms], v_sel)
    assert bool((ok[:, ms:] == 0).all())


# ---------------------------------------------------------------- reference-style unit tests
def test_async_prefill_statistics_in_one_transfer(rt):
    """a whole prefill without host synchronisation, statistics of all layers fetched once;
    the resulting compression_info / layer_states equal the synchronous per-layer call."""
    rtkv, ops = rt
    L = 4
    cfg = rtkv.CompressionConfig(num_hidden_layers=L, num_attention_heads=8, hidden_size=512, high_precision_bits=8,
                                 medium_precision_bits=4, low_precision_bits=2)
    sync, asyn = rtkv.RealTimePrefillCompressor(cfg), rtkv.RealTimePrefillCompressor(cfg)
    layers = [synth_layer(2, 8, 300, 512, seed=10 + l)[:3] for l in range(L)]
    ids = torch.zeros(2, 300, dtype=torch.long, device="cuda")
    pending = []
    for l, (K, V, A) in enumerate(layers):
        ok, ov = torch.zeros_like(K), torch.zeros_like(V)
        sel = asyn.compress_layer_kv_cache_async(K, V, A, l, ok, ov, zero_pad=True)
        pending.append((l, sel, ok, ov))
    got = asyn.collect_async(pending)
    for l, (K, V, A) in enumerate(layers):
        k, v, info = sync.compress_layer_kv_cache(K, V, A, ids, l)
        gk, gv, ginfo = got[l]
        assert torch.equal(gk, k) and torch.equal(gv, v)
        for key in ("compressed_shape", "compression_ratio", "memory_savings", "importance_stats", "precision_stats"):
            assert ginfo[key] == info[key], key
        np.testing.assert_array_equal(ginfo["quantization_info"]["bit_assignments"], info["quantization_info"]["bit_assignments"])
        assert torch.equal(ginfo["propagation_info"]["selection_mask"], info["propagation_info"]["selection_mask"])
        assert ginfo["propagation_info"]["selection_stats"] == info["propagation_info"]["selection_stats"]
        assert torch.equal(asyn.importance_tracker.layer_scores[l], sync.importance_tracker.layer_scores[l])
    a, b = asyn.get_overall_compression_stats(), sync.get_overall_compression_stats()
    for key in a:
        if "time" not in key:
            assert a[key] == b[key], key


def test_reference_style_component_tests(rt):
    """The assertions of the reference's tests/test_compression.py and test_quantization.py."""
    rtkv, ops = rt
    cfg = rtkv.CompressionConfig(model_name="meta-llama/Llama-2-7b-hf", num_hidden_layers=4, alpha=0.4, beta=0.3,
                                 gamma=0.3, theta_h=0.7, theta_m=0.3)
    attn = torch.softmax(torch.randn(1, 8, 10, 10), dim=-1)
    K, V = torch.randn(1, 10, 64), torch.randn(1, 10, 64)
    scorer = rtkv.PromptGuidedImportanceScorer(cfg)
    s = scorer.compute_importance_scores(attn, torch.arange(3), layer_idx=0)
    assert s.shape == (1, 10) and torch.isfinite(s).all() and s.device.type == "cpu"
    quant = rtkv.DynamicPrecisionQuantizer(cfg)
    labels, stats = quant.assign_precision_levels(torch.randn(1, 10))
    assert labels.shape == (1, 10) and labels.dtype == torch.long
    assert all(l in [0, 1, 2] for l in labels.unique().tolist())
    qk, qv, qi = quant.apply_mixed_precision_quantization(K, V, labels)
    assert qk.shape == K.shape and qv.shape == V.shape and torch.isfinite(qk).all()
    prop = rtkv.SelectiveTokenPropagator(cfg)
    sk, sv, ss, sl, info = prop.apply_token_selection(K, V, torch.randn(1, 10), torch.randint(0, 3, (1, 10)), layer_idx=0)
    assert sk.shape[1] <= 10 and sk.shape[1] == sv.shape[1]
    mem = quant.estimate_memory_savings(K, torch.randint(0, 3, (1, 10)))
    assert set(mem) == {"original_memory_mb", "compressed_memory_mb", "compression_ratio", "memory_savings",
                        "high_elements_ratio", "medium_elements_ratio", "low_elements_ratio"}
    costs = prop.compute_token_costs(torch.tensor([[0, 1, 2]]))
    assert costs.tolist() == [[0.5, 1.0, 2.0]]


def test_reference_style_scorer_tests(rt):
    """tests/test_importance_scoring.py (with the min-max assertion the code actually implements)."""
    rtkv, ops = rt
    cfg = rtkv.CompressionConfig(alpha=0.4, beta=0.3, gamma=0.3, num_hidden_layers=8)
    scorer = rtkv.PromptGuidedImportanceScorer(cfg)
    attn = torch.softmax(torch.randn(2, 8, 16, 16), dim=-1)
    pidx = torch.arange(4)
    agg = scorer.compute_attention_aggregation(attn, pidx, 0)
    assert agg.shape == (2, 16) and (agg >= 0).all() and torch.isfinite(agg).all()
    torch.testing.assert_close(agg, attn[:, :, :, pidx].mean(1).sum(-1), rtol=1e-5, atol=1e-7)
    nrm = scorer.normalize_attention_scores(agg, 0)
    assert nrm.shape == agg.shape and float(nrm.min()) == 0.0 and float(nrm.max()) == 1.0
    pb = scorer.compute_position_bias(50, torch.device("cpu"))
    assert pb.shape == (50,) and (pb >= 0).all() and (pb <= 1).all() and (pb[1:] >= pb[:-1]).all()
    assert abs(pb[0].item()) < 1e-6 and abs(pb[-1].item() - 1.0) < 1e-6
    cr = scorer.compute_context_relevance(20, 4, torch.device("cpu"))
    assert cr.shape == (20,) and (cr == cr[0]).all() and abs(cr[0].item() - 0.2) < 1e-6
    s0 = scorer.compute_importance_scores(attn, pidx, 0)
    s7 = scorer.compute_importance_scores(attn, pidx, 7)
    assert not torch.allclose(s0, s7, atol=1e-6)
    edge = scorer.compute_importance_scores(torch.softmax(torch.randn(1, 2, 2, 2), -1), torch.tensor([0]), 0)
    assert edge.shape == (1, 2) and torch.isfinite(edge).all()
    tr = rtkv.LayerWiseImportanceTracker(cfg)
    for l in range(3):
        tr.update_scores(l, attn[:1, :, :12, :12].contiguous(), torch.arange(3))
    cum = tr.get_cumulative_scores(2)
    assert torch.allclose(cum, (tr.layer_scores[0] + tr.layer_scores[1] + tr.layer_scores[2]) / 3)
    with pytest.raises(IndexError):
        scorer.compute_importance_scores(attn, pidx, 8)         # layer_idx >= len(layer_weights)


def test_llama_hook_prefill_on_gpu(rt):
    """The hook contract of modified_llama.py:96-152 on a tiny random Llama, model level: every layer is compressed,
    statistics are attached, the cache entry is the compressed K/V.  (The numerics of the block are checked against
    a hand recomputation in test_llama_hook_block_equals_hand_recomputation.)"""
    rtkv, ops = rt
    from rtkv.llama_hooks import CompressedLlamaForCausalLM
    from transformers import LlamaConfig
    torch.manual_seed(0)
    hf = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=8,
                     num_key_value_heads=8, vocab_size=211, max_position_embeddings=256)
    model = CompressedLlamaForCausalLM(hf).cuda().eval()
    cfg = rtkv.CompressionConfig(num_hidden_layers=3, num_attention_heads=8, hidden_size=128, high_precision_bits=8,
                                 medium_precision_bits=8, low_precision_bits=8)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    model.set_compressor(comp)
    ids = torch.randint(0, 211, (2, 60), device="cuda")
    with torch.no_grad():
        out = model(input_ids=ids, use_cache=True)
    assert torch.isfinite(out.logits).all() and out.logits.shape == (2, 60, 211)
    assert sorted(comp.layer_states) == [0, 1, 2]
    # all-8-bit: cost 1.0 per token -> exactly floor(ratio*S) tokens survive
    assert [comp.layer_states[l]["compressed_shape"][1] for l in range(3)] == [36, 36, 24]   # L=3: int(0.3*3)=0 early layers, int(0.7*3)=2
    assert out.compression_info["total_layers_processed"] == 3
    assert abs(out.compression_info["cumulative_compression"] - 24 / 60) < 1e-12
    kept = out.past_key_values.layers[2].keys.shape[2]
    assert kept == 24                                            # the cache entry is the compressed K/V
    model.reset_compression_state()
    assert comp.layer_states == {}


@pytest.mark.parametrize("bits,prunes", [((8, 8, 8), True), ((2, 2, 2), False)])
def test_llama_hook_block_equals_hand_recomputation(rt, bits, prunes):
    """Row a19: the hook block of CompressedLlamaAttention.forward (modified_llama.py:96-152) recomputed by hand in
    plain torch, with the ORACLE doing the compression step: eager softmax(QK^T/sqrt(d) + mask), compressor fed
    K/V as [B,S,H*D] and the probabilities, then either attention recomputed over the kept tokens with the additive
    mask truncated to its first S' columns (:125-140, S' != S) or attn_weights @ V_quant (S' == S), o_proj.  Also the
    downgrade when the compressor raises (:144-149)."""
    rtkv, ops = rt
    from rtkv.llama_hooks import CompressedLlamaForCausalLM
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    torch.manual_seed(1)
    Hh, D, S, B, Lh = 8, 16, 60, 2, 3
    hf = LlamaConfig(hidden_size=Hh * D, intermediate_size=256, num_hidden_layers=Lh, num_attention_heads=Hh,
                     num_key_value_heads=Hh, vocab_size=211, max_position_embeddings=256)
    model = CompressedLlamaForCausalLM(hf).cuda().eval()
    cfg = rtkv.CompressionConfig(num_hidden_layers=Lh, num_attention_heads=Hh, hidden_size=Hh * D, high_precision_bits=bits[2],
                                 medium_precision_bits=bits[1], low_precision_bits=bits[0])
    comp = rtkv.RealTimePrefillCompressor(cfg)
    model.set_compressor(comp)
    layer = 2
    attn = model.model.layers[layer].self_attn
    x = torch.randn(B, S, Hh * D, device="cuda")
    pos = torch.arange(S, device="cuda")[None].expand(B, -1)
    cos, sin = model.model.rotary_emb(x, pos)
    mask = torch.full((S, S), torch.finfo(torch.float32).min, device="cuda").triu(1)[None, None].expand(B, 1, S, S)
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    with torch.no_grad():
        got, got_w = attn(x, position_embeddings=(cos, sin), attention_mask=mask, input_ids=ids)
        # ---- by hand ----
        shp = (B, S, Hh, D)
        q = torch.nn.functional.linear(x, attn.q_proj.weight).view(shp).transpose(1, 2)
        k = torch.nn.functional.linear(x, attn.k_proj.weight).view(shp).transpose(1, 2)
        v = torch.nn.functional.linear(x, attn.v_proj.weight).view(shp).transpose(1, 2)
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        w = torch.softmax(q @ k.transpose(2, 3) / D ** 0.5 + mask, dim=-1, dtype=torch.float32)
        assert torch.equal(got_w, w)                      # the probabilities the reference hands to the compressor (:88-94)
        k_flat = k.transpose(1, 2).contiguous().view(B, S, Hh * D)
        v_flat = v.transpose(1, 2).contiguous().view(B, S, Hh * D)
        P = max(1, min(S // 5, 128))
        o = oracle.compress_layer(host(k_flat), host(v_flat), oracle.F32, host(w[..., :P].contiguous()), oracle.F32, P,
                                  cfg.alpha, cfg.beta, cfg.gamma, cfg.layer_weights[layer], cfg.theta_h, cfg.theta_m, bits,
                                  comp.propagator.get_layer_propagation_ratio(layer))
        ck, cv = torch.from_numpy(o["k_sel"]).cuda(), torch.from_numpy(o["v_sel"]).cuda()
        s_kept = ck.shape[1]
        assert (s_kept != S) == prunes
        ck = ck.view(B, s_kept, Hh, D).transpose(1, 2)
        cv = cv.view(B, s_kept, Hh, D).transpose(1, 2)
        if s_kept != S:
            w2 = torch.softmax(q @ ck.transpose(2, 3) / D ** 0.5 + mask[..., :s_kept], dim=-1, dtype=torch.float32)
            ctx = w2 @ cv
        else:
            ctx = w @ cv
        want = torch.nn.functional.linear(ctx.transpose(1, 2).contiguous().reshape(B, S, -1), attn.o_proj.weight)
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
        assert comp.layer_states[layer]["compressed_shape"][1] == s_kept
        # ---- downgrade: a failing compressor leaves the plain attention result (:144-149) ----

        class Boom:
            def compress_layer_kv_cache(self, *a, **k):
                raise RuntimeError("boom")
        attn.set_compressor(Boom())
        got2, _ = attn(x, position_embeddings=(cos, sin), attention_mask=mask, input_ids=ids)
        plain = torch.nn.functional.linear((w @ v).transpose(1, 2).contiguous().reshape(B, S, -1), attn.o_proj.weight)
        torch.testing.assert_close(got2, plain, rtol=1e-5, atol=1e-6)
        # training mode / decode step (q_len == 1): the block is skipped (:102)
        attn.set_compressor(comp)
        comp.reset_compression_state()
        attn.train()
        attn(x, position_embeddings=(cos, sin), attention_mask=mask, input_ids=ids)
        assert comp.layer_states == {}


def test_precomputed_prompt_mass_input(rt):
    """compress_layer_kv_cache(prompt_mass=...) skips K1a: same result as handing in the attention tensor."""
    rtkv, ops = rt
    K, V, A, P = synth_layer(2, 8, 300, 512, seed=4)
    cfg = rtkv.CompressionConfig(num_hidden_layers=8, num_attention_heads=8, hidden_size=512, high_precision_bits=8,
                                 medium_precision_bits=4, low_precision_bits=2)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(2, 300, dtype=torch.long, device="cuda")
    k0, v0, i0 = comp.compress_layer_kv_cache(K, V, A, ids, 5)
    k1, v1, i1 = comp.compress_layer_kv_cache(K, V, None, ids, 5, prompt_mass=ops.prompt_mass(A, prompt_len=P))
    assert torch.equal(k0, k1) and torch.equal(v0, v1)
    assert torch.equal(i0["propagation_info"]["selection_mask"], i1["propagation_info"]["selection_mask"])
    with pytest.raises(ValueError):
        comp.compress_layer_kv_cache(K, V, None, ids, 5, prompt_mass=torch.zeros(2, 299, device="cuda"))


def test_llama_hook_fused_scoring_no_attention_matrix(rt):
    """the hook with librtkv's tcgen05 prefill attention (output + LSE + fused prompt mass, and the
    attention over the kept tokens) instead of a materialised softmax(QK^T).  Same weights as the eager hook: same kept counts, (nearly) the same layer-0
    selection, close logits; attention weights are not returned."""
    rtkv, ops = rt
    from rtkv.llama_hooks import CompressedLlamaForCausalLM
    from transformers import LlamaConfig
    torch.manual_seed(1)
    hf = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                     num_key_value_heads=2, vocab_size=211, max_position_embeddings=1024)
    eager = CompressedLlamaForCausalLM(hf).cuda().to(torch.bfloat16).eval()
    fused = CompressedLlamaForCausalLM(hf).cuda().to(torch.bfloat16).eval()
    fused.load_state_dict(eager.state_dict())
    kw = dict(num_hidden_layers=3, num_attention_heads=4, hidden_size=512, high_precision_bits=8,
              medium_precision_bits=8, low_precision_bits=8)
    c_eager, c_fused = rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig(**kw)), \
        rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig(**kw))
    eager.set_compressor(c_eager)
    fused.set_compressor(c_fused, fused_scoring=True)
    ids = torch.randint(0, 211, (2, 640), device="cuda")
    with torch.no_grad():
        o_e = eager(input_ids=ids, use_cache=True)
        o_f = fused(input_ids=ids, use_cache=True, output_attentions=False)
    assert torch.isfinite(o_f.logits).all()
    assert sorted(c_fused.layer_states) == [0, 1, 2]
    shapes = lambda c: [c.layer_states[l]["compressed_shape"][1] for l in range(3)]
    assert shapes(c_fused) == shapes(c_eager) == [384, 384, 256]
    m_e = c_eager.layer_states[0]["propagation_info"]["selection_mask"]
    m_f = c_fused.layer_states[0]["propagation_info"]["selection_mask"]
    overlap = float((m_e & m_f).sum()) / float(m_e.sum())
    assert overlap > 0.8, overlap          # the eager bf16 path rounds the mass to bf16 like the reference does
    assert o_f.past_key_values.layers[2].keys.shape[2] == 256


def test_llama_hook_fused_scoring_needs_a_causal_mask(rt):
    """The fused branch only runs for a None / purely causal mask; a padding mask (or output_attentions=True) takes the
    eager branch, which returns the attention weights."""
    rtkv, ops = rt
    from rtkv.llama_hooks import CompressedLlamaForCausalLM, _is_causal_mask
    from transformers import LlamaConfig
    torch.manual_seed(3)
    hf = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=2, vocab_size=97, max_position_embeddings=512)
    model = CompressedLlamaForCausalLM(hf).cuda().to(torch.bfloat16).eval()
    cfg = rtkv.CompressionConfig(num_hidden_layers=2, num_attention_heads=2, hidden_size=256, high_precision_bits=8,
                                 medium_precision_bits=8, low_precision_bits=8)
    model.set_compressor(rtkv.RealTimePrefillCompressor(cfg), fused_scoring=True)
    attn = model.model.layers[0].self_attn
    B, S = 2, 200
    x = torch.randn(B, S, 256, device="cuda", dtype=torch.bfloat16)
    pos = torch.arange(S, device="cuda")[None].expand(B, -1)
    pe = model.model.rotary_emb(x, pos)
    neg = torch.finfo(torch.bfloat16).min
    causal = torch.full((S, S), neg, device="cuda", dtype=torch.bfloat16).triu(1)[None, None].expand(B, 1, S, S).contiguous()
    padded = causal.clone()
    padded[1, :, :, :7] = neg                       # left padding of the second sequence
    assert _is_causal_mask(None, S, S) and _is_causal_mask(causal, S, S) and not _is_causal_mask(padded, S, S)
    with torch.no_grad():
        _, w_none = attn(x, position_embeddings=pe, attention_mask=None)
        _, w_causal = attn(x, position_embeddings=pe, attention_mask=causal)
        _, w_padded = attn(x, position_embeddings=pe, attention_mask=padded)
        _, w_asked = attn(x, position_embeddings=pe, attention_mask=causal, output_attentions=True)
    assert w_none is None and w_causal is None               # fused branch: no probabilities exist
    assert w_padded is not None and w_padded.shape == (B, 2, S, S) and w_asked is not None


def test_llama_hook_fused_scoring_32k_tokens(rt):
    """BASELINE configs[2] shape through the model hook: 32 heads x 128, 32768 tokens.  The eager formulation of the
    reference needs a [1,32,32768,32768] probability tensor (64 GiB in bf16, twice that for the fp32 softmax); the fused
    hook never builds it."""
    rtkv, ops = rt
    from rtkv.llama_hooks import CompressedLlamaForCausalLM
    from transformers import LlamaConfig
    torch.manual_seed(2)
    S = 32768
    hf = LlamaConfig(hidden_size=4096, intermediate_size=512, num_hidden_layers=2, num_attention_heads=32,
                     num_key_value_heads=32, vocab_size=64, max_position_embeddings=S)
    model = CompressedLlamaForCausalLM(hf).cuda().to(torch.bfloat16).eval()
    cfg = rtkv.CompressionConfig(num_hidden_layers=2, num_attention_heads=32, hidden_size=4096, high_precision_bits=8,
                                 medium_precision_bits=8, low_precision_bits=8)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    model.set_compressor(comp, fused_scoring=True)
    ids = torch.randint(0, 64, (1, S), device="cuda")
    with torch.no_grad():
        x = model.model.embed_tokens(ids)
        pos = torch.arange(S, device="cuda")[None]
        pe = model.model.rotary_emb(x, pos)
        for layer in model.model.layers:            # decoder layers directly: HF's eager mask builder would allocate S x S
            x = layer(x, position_embeddings=pe, attention_mask=None, input_ids=ids)
            x = x[0] if isinstance(x, tuple) else x
    assert torch.isfinite(x).all()
    assert [comp.layer_states[l]["compressed_shape"][1] for l in range(2)] == [int(0.6 * S), int(0.4 * S)]
    assert torch.cuda.max_memory_allocated() < 24 << 30


# ---------------------------------------------------------------- K1b fused Q.K scoring (tcgen05)
def qk_inputs(B, H, S, P, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = 128
    q = (torch.randn(B, H, S, D, generator=g, device="cuda") * 1.2).to(dtype)
    k = (torch.randn(B, H, S, D, generator=g, device="cuda") * 1.2).to(dtype)
    scale = D ** -0.5
    logits = (q.float() @ k.float().transpose(-1, -2)) * scale
    logits = logits + torch.full((S, S), float("-inf"), device="cuda").triu(1)
    lse = torch.logsumexp(logits, dim=-1)
    probs = torch.softmax(logits, dim=-1)[..., :P].contiguous()
    return q, k[:, :, :P].contiguous(), lse, scale, probs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,S,P", [(1, 8, 256, 51), (2, 16, 640, 128), (1, 32, 1000, 128), (1, 8, 130, 26),
                                      (1, 32, 4096, 128), (1, 4, 8200, 128)])
def test_qk_scoring_matches_restatement(rt, dtype, B, H, S, P):
    """tensor-core scoring vs the CPU restatement (double accumulation) and vs the probability-slice
    kernel on the same attention: north-star tolerance for half inputs is 1e-3 relative."""
    _, ops = rt
    q, kp, lse, scale, probs = qk_inputs(B, H, S, P, dtype, seed=S + H)
    parts = ops.prompt_mass_qk_partials(q, kp, lse, scale, causal=True)            # per-head sums [B,H,S]
    dt = oracle.BF16 if dtype == torch.bfloat16 else oracle.F16
    want = oracle.prompt_mass_qk(oracle.from_torch(q)[0], oracle.from_torch(kp)[0], lse.cpu().numpy(), dt, scale, True, 1)
    np.testing.assert_allclose(parts.transpose(1, 2).cpu().numpy(), want, rtol=2e-4, atol=1e-6)
    mass = ops.prompt_mass_qk(q, kp, lse, scale)
    np.testing.assert_allclose(mass.cpu().numpy(), oracle.mass_from_partials(want, H), rtol=2e-4, atol=1e-6)
    via_probs = ops.prompt_mass(probs, prompt_len=P)
    np.testing.assert_allclose(mass.cpu().numpy(), via_probs.cpu().numpy(), rtol=1e-3, atol=1e-6)


def test_qk_scoring_rejects_unsupported(rt):
    _, ops = rt
    q = torch.zeros(1, 8, 128, 128, device="cuda")
    with pytest.raises(RuntimeError, match="bf16 or f16"):
        ops.prompt_mass_qk_partials(q, q, torch.zeros(1, 8, 128, device="cuda"), 0.1)
    q64 = torch.zeros(1, 8, 128, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="head_dim == 128"):
        ops.prompt_mass_qk_partials(q64, q64, torch.zeros(1, 8, 128, device="cuda"), 0.1)
