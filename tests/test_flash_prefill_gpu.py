"""rtkv_flash_prefill (tcgen05 + TMA prefill attention with LSE and fused prompt-column mass) against a plain PyTorch
fp32 reference of the same op: the reference's eager formulation softmax(Q K^T / sqrt(d) + mask) V of
src/models/modified_llama.py:88-94 (and :125-142 with the mask truncated to the first S' columns when the keys are the
S' kept tokens).  Floating point kernel: tolerances are written next to every comparison (north star: 1e-3 relative for
bf16 scales; the attention output itself is a bf16/f16 tensor, i.e. 2^-8 / 2^-11 relative per element)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv  # noqa: F401
    from rtkv import ops
    return ops


def reference(q, k, v, scale, causal, P):
    """fp32 eager attention on the (already rounded) half-precision inputs."""
    B, H, Sq, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    qf, kf, vf = q.float(), k.float().repeat_interleave(H // Hkv, dim=1), v.float().repeat_interleave(H // Hkv, dim=1)
    logits = qf @ kf.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]
        j = torch.arange(Skv, device=q.device)[None, :]
        logits = logits.masked_fill(j > i, float("-inf"))
    lse = torch.logsumexp(logits, dim=-1)
    probs = torch.softmax(logits, dim=-1)
    out = probs @ vf
    hs = probs[..., :P].sum(-1) if P > 0 else None
    return out, lse, hs


CASES = [
    # B, H, Hkv, Sq, Skv, causal, P, dtype, qscale
    (1, 2, 2, 128, 128, True, 128, torch.bfloat16, 1.0),
    (2, 4, 2, 300, 300, True, 60, torch.bfloat16, 1.0),
    (1, 2, 2, 1000, 1000, True, 128, torch.float16, 1.0),
    (1, 2, 1, 256, 200, False, 0, torch.bfloat16, 1.0),
    (1, 4, 4, 700, 420, True, 0, torch.bfloat16, 1.0),          # keys = kept tokens, top-left aligned (truncated) causal mask
    (1, 2, 2, 640, 640, True, 128, torch.bfloat16, 6.0),         # sharp logits: the lazy rescale of O fires
    (1, 3, 3, 2048, 2048, True, 128, torch.bfloat16, 1.0),
    (1, 2, 2, 77, 77, True, 15, torch.float16, 1.0),
]


@pytest.fixture(params=[0, 1, 2, 3, 4], ids=["halves-exchange-max", "independent-key-halves", "two-tiles-P-in-TMEM", "alternating-key-tiles", "alternating-tiles-one-accumulator"])
def variant(request, ops):
    """The five kernels behind rtkv_flash_prefill: 0 = one query tile per CTA, the two softmax warpgroups agree on the row max
    every tile, 1 = same layout, independent online softmaxes per key half, 2 = two query tiles per CTA with P in tensor
    memory, 3 = each softmax warpgroup owns every other key tile (own accumulator, merged in the epilogue), 4 = alternating key
    tiles with the running max handed from tile to tile, one accumulator, P in tensor memory, three K/V stages."""
    ops.lib.rtkv_debug_flash_variant(request.param)
    yield request.param
    ops.lib.rtkv_debug_flash_variant(-1)


@pytest.mark.parametrize("B,H,Hkv,Sq,Skv,causal,P,dtype,qscale", CASES)
def test_flash_prefill_matches_fp32_reference(ops, variant, B, H, Hkv, Sq, Skv, causal, P, dtype, qscale):
    g = torch.Generator(device="cuda").manual_seed(Sq * 7 + Skv)
    D = 128
    q = (torch.randn(B, H, Sq, D, generator=g, device="cuda") * qscale).to(dtype)
    k = torch.randn(B, Hkv, Skv, D, generator=g, device="cuda").to(dtype)
    v = torch.randn(B, Hkv, Skv, D, generator=g, device="cuda").to(dtype)
    scale = 1.0 / math.sqrt(D)
    out, lse, hs = ops.flash_prefill(q, k, v, scale, causal=causal, prompt_len=P)
    want, want_lse, want_hs = reference(q, k, v, scale, causal, P)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    # output: a half-precision tensor; P is rounded to the input dtype before the second product -> a few eps
    torch.testing.assert_close(out.float(), want, rtol=4 * eps, atol=4 * eps)
    torch.testing.assert_close(lse, want_lse, rtol=0, atol=2e-3)          # |lse| ~ 10: ~2e-4 relative
    if P > 0:
        torch.testing.assert_close(hs, want_hs, rtol=2e-3, atol=1e-6)     # ex2.approx + fp32 sums


def test_flash_prefill_strided_inputs_and_bshd_output(ops, variant):
    """Inputs as the transposed view of [B,S,H,D] projections (what the model hook holds), output written as [B,S,H,D]."""
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, S, D = 2, 4, 333, 128
    q = torch.randn(B, S, H, D, generator=g, device="cuda").to(torch.bfloat16).transpose(1, 2)
    k = torch.randn(B, S, H, D, generator=g, device="cuda").to(torch.bfloat16).transpose(1, 2)
    v = torch.randn(B, S, H, D, generator=g, device="cuda").to(torch.bfloat16).transpose(1, 2)
    assert not q.is_contiguous()
    out, lse, hs = ops.flash_prefill(q, k, v, causal=True, prompt_len=66, out_layout="bshd")
    want, want_lse, want_hs = reference(q, k, v, D ** -0.5, True, 66)
    assert out.shape == (B, S, H, D) and out.is_contiguous()
    torch.testing.assert_close(out.float(), want.transpose(1, 2), rtol=2.0 ** -6, atol=2.0 ** -6)
    torch.testing.assert_close(lse, want_lse, rtol=0, atol=2e-3)
    torch.testing.assert_close(hs, want_hs, rtol=2e-3, atol=1e-6)


def test_flash_prefill_prompt_mass_feeds_the_compressor(ops, variant):
    """head_sums -> rtkv_mass_from_head_sums == the mass K1a computes from materialised probabilities, within the
    tolerance the scoring kernel is held to (2e-4 relative in test_qk_scoring_matches_restatement)."""
    import rtkv
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, S, D = 1, 8, 900, 128
    q = torch.randn(B, H, S, D, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, H, S, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, H, S, D, generator=g, device="cuda").to(torch.bfloat16)
    P = rtkv.RealTimePrefillCompressor.prompt_length(S)
    _, _, hs = ops.flash_prefill(q, k, v, causal=True, prompt_len=P)
    mass = torch.empty(B, S, dtype=torch.float32, device="cuda")
    ops.check(ops.lib.rtkv_mass_from_head_sums(ops._ptr(hs), B, H, S, ops._ptr(mass), ops._stream()))
    logits = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(D)
    logits = logits + torch.full((S, S), float("-inf"), device="cuda").triu(1)
    A = torch.softmax(logits, -1)[..., :P].contiguous()
    want = ops.prompt_mass(A, prompt_len=P)
    torch.testing.assert_close(mass, want, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("every", [2, 3, 4])
def test_flash_prefill_polynomial_exponentials(ops, every):
    """rtkv_debug_flash_poly: a share of the exponentials of unmasked tiles on the FMA pipe (Cody-Waite split + degree-4
    polynomial, relative error < 1e-4 — below the bf16 rounding of P).  Off by default (slower); same tolerances."""
    ops.lib.rtkv_debug_flash_poly(every)
    try:
        test_flash_prefill_matches_fp32_reference(ops, 0, 1, 3, 3, 2048, 2048, True, 128, torch.bfloat16, 1.0)
        test_flash_prefill_matches_fp32_reference(ops, 0, 1, 2, 2, 1000, 1000, True, 128, torch.float16, 1.0)
    finally:
        ops.lib.rtkv_debug_flash_poly(0)


@pytest.mark.parametrize("variant_id", [0, 4])
def test_flash_prefill_phase_profiler_leaves_results_alone(ops, variant_id):
    """rtkv_debug_flash_profile: the instrumented build (cycle counters per phase of CTA 0's tile loop, tools/profile_flash_phases.py)
    computes the same output as the production build and fills its counters."""
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v = (torch.randn(1, 2, 1024, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
    ops.lib.rtkv_debug_flash_variant(variant_id)
    try:
        want = ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
        buf = torch.zeros(512, dtype=torch.int64, device="cuda")
        ops.lib.rtkv_debug_flash_profile(buf.data_ptr())
        try:
            got = ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
            torch.cuda.synchronize()
        finally:
            ops.lib.rtkv_debug_flash_profile(None)
    finally:
        ops.lib.rtkv_debug_flash_variant(-1)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    c = buf.cpu()
    assert int(c[8]) > 0 and int(c[6]) > 0 and int(c[33]) > 0          # tiles, cycles in the exponentials, MMA thread loop


def test_tensor_pipe_rate_probe(ops):
    """rtkv_debug_mma_rate (tools/mma_rate.py): runs, and with precomputed descriptors an MMA is issued no slower than when the
    descriptors are rebuilt per instruction."""
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    res = {}
    for mode in (0, 6):
        ops.check(ops.lib.rtkv_debug_mma_rate(mode, 512, 4, buf.data_ptr(), None))
        torch.cuda.synchronize()
        res[mode] = buf.cpu().tolist()[:2]
        assert res[mode][0] > 0 and res[mode][1] >= res[mode][0]
    assert res[6][1] <= res[0][1]
    with pytest.raises(RuntimeError):
        ops.check(ops.lib.rtkv_debug_mma_rate(9, 512, 4, buf.data_ptr(), None))


def test_flash_prefill_rejects_unsupported(ops):
    q = torch.randn(1, 2, 64, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.flash_prefill(q, q, q)
    q = torch.randn(1, 2, 64, 128, device="cuda", dtype=torch.float32)
    with pytest.raises((RuntimeError, TypeError)):
        ops.flash_prefill(q, q, q)
