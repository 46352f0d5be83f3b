"""Decode-step attention over packed rows (rtkv_packed_decode_attention): a project-defined
consumer of the compressed cache with no reference counterpart.  Checked against plain fp64 attention over the
dequantised rows (PackedKV.dequantize(), itself bit-checked against the oracle); tolerance 2e-5 relative — the kernel
folds the dequantisation into the dot products, so only the fp32 summation order differs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built_lib):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv
    from rtkv import ops
    return rtkv, ops


def reference_attention(q, dk, dv, counts, hkv):
    B, H, D = q.shape
    out = torch.zeros(B, H, D, dtype=torch.float64, device=q.device)
    lse = torch.zeros(B, H, dtype=torch.float64, device=q.device)
    for b in range(B):
        c = int(counts[b])
        k = dk[b, :c].double().view(c, hkv, D)
        v = dv[b, :c].double().view(c, hkv, D)
        for h in range(H):
            hk = h // (H // hkv)
            s = (k[:, hk] @ q[b, h].double()) * D ** -0.5
            p = torch.softmax(s, dim=0)
            out[b, h] = p @ v[:, hk]
            lse[b, h] = torch.logsumexp(s, dim=0)
    return out, lse


@pytest.mark.parametrize("bits", [(2, 4, 8), (4, 8, 8), (1, 2, 4)])
@pytest.mark.parametrize("B,Hq,S,HD", [(2, 8, 300, 512), (1, 32, 2000, 4096), (3, 4, 70, 128)])
def test_attention_over_packed_rows(rt, bits, B, Hq, S, HD):
    rtkv, ops = rt
    from test_parity_gpu import synth_layer
    K, V, A, P = synth_layer(B, 8, S, HD, seed=S + bits[0])
    cfg = rtkv.CompressionConfig(num_hidden_layers=8, num_attention_heads=8, hidden_size=HD, low_precision_bits=bits[0],
                                 medium_precision_bits=bits[1], high_precision_bits=bits[2])
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    packed = comp.compress_layer_kv_cache_packed(K, V, A, ids, 6)
    hkv = HD // 128
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(B, Hq, 128, generator=g, device="cuda") * 2.0
    dk, dv = packed.dequantize()
    want, want_lse = reference_attention(q, dk, dv, packed.rows["n_selected"], hkv)
    for splits in (None, 1, 3):
        if splits is None:
            got, lse = packed.attend(q, hkv, want_lse=True)
        else:
            got, lse = ops.packed_decode_attention(q, packed.keys["codes"], packed.keys["qparams"], packed.values["codes"],
                                                   packed.values["qparams"], packed._row_bits, packed.sel, hkv,
                                                   splits=splits, want_lse=True)
        scale = want.abs().max()
        assert float((got.double() - want).abs().max() / scale) < 2e-5
        assert float((lse.double() - want_lse).abs().max()) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Hq,S,HD,Sq", [(1, 8, 2000, 1024, 2000), (2, 4, 600, 512, 333)])
def test_prefill_attention_over_packed_rows(rt, dtype, B, Hq, S, HD, Sq):
    """PackedKV.attend_prefill: queries of a prefill against the packed kept rows (unpack + dequantise, then the tcgen05 attention
    kernel) == fp32 eager attention over dequantize() with the reference's truncated causal mask (modified_llama.py:125-140)."""
    rtkv, ops = rt
    from test_parity_gpu import synth_layer
    K, V, A, P = synth_layer(B, 8, S, HD, seed=S + 3)
    cfg = rtkv.CompressionConfig(num_hidden_layers=8, num_attention_heads=8, hidden_size=HD, low_precision_bits=4,
                                 medium_precision_bits=8, high_precision_bits=8)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    packed = comp.compress_layer_kv_cache_packed(K, V, A, torch.zeros(B, S, dtype=torch.long, device="cuda"), 6)
    hkv = HD // 128
    g = torch.Generator(device="cuda").manual_seed(9)
    q = torch.randn(B, Hq, Sq, 128, generator=g, device="cuda").to(dtype)
    out, lse = packed.attend_prefill(q, hkv, want_lse=True)
    dk, dv = packed.dequantize()
    kept = dk.shape[1]
    assert kept == packed.max_selected and kept < S                      # the budget binds: S' < S
    k4 = dk.to(dtype).float().view(B, kept, hkv, 128).transpose(1, 2).repeat_interleave(Hq // hkv, dim=1)
    v4 = dv.to(dtype).float().view(B, kept, hkv, 128).transpose(1, 2).repeat_interleave(Hq // hkv, dim=1)
    logits = (q.float() @ k4.transpose(-1, -2)) * 128 ** -0.5
    i = torch.arange(Sq, device="cuda")[:, None]
    j = torch.arange(kept, device="cuda")[None, :]
    logits = logits.masked_fill(j > i, float("-inf"))
    want = torch.softmax(logits, -1) @ v4
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    torch.testing.assert_close(out.float(), want, rtol=4 * eps, atol=4 * eps)
    torch.testing.assert_close(lse, torch.logsumexp(logits, -1), rtol=0, atol=2e-3)


def test_rejects_unsupported(rt):
    rtkv, ops = rt
    from test_parity_gpu import synth_layer
    K, V, A, P = synth_layer(1, 8, 100, 512, seed=1)
    cfg = rtkv.CompressionConfig(num_hidden_layers=8, num_attention_heads=8, hidden_size=512, low_precision_bits=2,
                                 medium_precision_bits=4, high_precision_bits=8, value_quant_mode="token_group")
    comp = rtkv.RealTimePrefillCompressor(cfg)
    packed = comp.compress_layer_kv_cache_packed(K, V, A, torch.zeros(1, 100, dtype=torch.long, device="cuda"), 0)
    with pytest.raises(NotImplementedError):
        packed.attend(torch.randn(1, 8, 128, device="cuda"))
    with pytest.raises(RuntimeError, match="head dimension"):
        ops.packed_decode_attention(torch.randn(1, 8, 64, device="cuda"), packed.keys["codes"], packed.keys["qparams"],
                                    packed.values["codes"], packed.values["qparams"],
                                    torch.zeros(1, packed.max_selected, dtype=torch.uint8, device="cuda"), packed.sel, 4)
