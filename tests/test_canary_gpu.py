"""Out-of-bounds write check without a sanitizer: every CUDA buffer the rtkv wrappers allocate during this test is carved out
of a larger allocation with 512 poisoned bytes on both sides; after running the kernels on awkward shapes the guard bands
must be untouched.  (compute-sanitizer is not available on the GPU pool.)"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PAD, POISON = 512, 0xA5


class GuardedTorch:
    """Stands in for the `torch` module inside rtkv.ops: empty()/zeros() on a CUDA device return guarded views."""

    def __init__(self, real):
        self._real, self.registry = real, []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _guarded(self, shape, dtype, device, fill):
        real = self._real
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, real.Size)):
            shape = tuple(shape[0])
        dtype = dtype or real.float32
        nbytes = int(math.prod(shape)) * real.empty(0, dtype=dtype).element_size()
        body = (nbytes + 15) // 16 * 16
        raw = real.full((body + 2 * PAD,), POISON, dtype=real.uint8, device=device)
        view = raw[PAD: PAD + nbytes].view(dtype).view(shape)
        if fill is not None:
            view.fill_(fill)
        self.registry.append((raw, nbytes))
        return view

    def empty(self, *shape, dtype=None, device=None, **kw):
        if device is None or self._real.device(device).type != "cuda":
            return self._real.empty(*shape, dtype=dtype, device=device, **kw)
        return self._guarded(shape, dtype, device, None)

    def zeros(self, *shape, dtype=None, device=None, **kw):
        if device is None or self._real.device(device).type != "cuda":
            return self._real.zeros(*shape, dtype=dtype, device=device, **kw)
        return self._guarded(shape, dtype, device, 0)

    def check(self):
        assert self.registry, "no guarded allocation was made"
        for raw, nbytes in self.registry:
            front, back = raw[:PAD], raw[PAD + nbytes:]
            assert bool((front == POISON).all()), f"write BEFORE a {nbytes}-byte buffer"
            assert bool((back == POISON).all()), f"write AFTER a {nbytes}-byte buffer"
        n = len(self.registry)
        self.registry.clear()
        return n


@pytest.fixture()
def guarded_ops(built_lib, monkeypatch):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv
    from rtkv import ops
    proxy = GuardedTorch(torch)
    monkeypatch.setattr(ops, "torch", proxy)
    return rtkv, ops, proxy


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,S,HD", [(2, 8, 333, 264), (1, 3, 70, 100), (3, 4, 1030, 512), (1, 40, 2100, 5120)])
def test_guard_bands_stay_intact(guarded_ops, dtype, B, H, S, HD):
    rtkv, ops, proxy = guarded_ops
    g = torch.Generator(device="cuda").manual_seed(S)
    P = max(1, min(S // 5, 128))
    A = torch.softmax(torch.randn(B, H, S, P + 3, generator=g, device="cuda") * 3, dim=-1).to(dtype)
    K = torch.randn(B, S, HD, generator=g, device="cuda").to(dtype)
    V = (torch.randn(B, S, HD, generator=g, device="cuda") * 3 + 1).to(dtype)
    cfg = rtkv.CompressionConfig(num_hidden_layers=8, num_attention_heads=H, hidden_size=HD, high_precision_bits=8,
                                 medium_precision_bits=4, low_precision_bits=2, early_layer_ratio=0.37)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    params, _ = comp.layer_params(1, S, ops.dtype_code(A))
    mass = ops.prompt_mass(A, prompt_len=P)
    ops.normalize_mass(mass)
    ops.position_bias(S, K.device)
    sel = ops.score_select(params, B, S, K.device, mass=mass)
    layer, rows = sel.read_stats()
    ms = int(layer["max_selected"])
    bits = (2, 4, 8)
    ops.quant_gather(K, V, sel, sel.labels, bits, ms)
    ops.quant_gather(K, None, sel, None, None, ms)                      # plain gather
    ops.gather_scalars(sel, ms)
    ops.row_minmax(K, V, sel, ms, negate_max=True)
    variants = [ops.quant_options("token", "token", outlier_tau=5.0)]
    if HD % 8 == 0:
        variants += [ops.quant_options("token_group", "token_group", group_size=8),
                     ops.quant_options("token_group", "channel", group_size=8, token_group=7, outlier_tau=4.0),
                     ops.quant_options("channel", "channel", token_group=32),
                     ops.quant_options("channel", "token", token_group=16)]
    for opt in variants:
        ops.quant_gather_ex(K, V, sel, sel.labels, bits, ms, opt, want_counts=True)
    codes, qp = ops.quant_pack(K, sel, sel.labels, bits, ms)
    ops.unpack_dequant(codes, qp, dtype, S, HD, sel, sel.labels, bits)
    if HD % 8 == 0:
        for opt, cap in ((variants[0], 5), (variants[1], 0), (variants[3], 0)):
            pk = ops.quant_pack_ex(V, sel, sel.labels, bits, ms, opt, tensor=0, outlier_cap=cap)
            ops.unpack_dequant_ex(pk, dtype, S, HD, sel, sel.labels, bits, opt, tensor=0)
    if HD % 128 == 0:
        packed = comp.compress_layer_kv_cache_packed(K, V, A, torch.zeros(B, S, dtype=torch.long, device="cuda"), 1)
        q = torch.randn(B, 2 * (HD // 128), 128, generator=g, device="cuda")
        packed.attend(q, want_lse=True)
        ops.packed_decode_attention(q, packed.keys["codes"], packed.keys["qparams"], packed.values["codes"],
                                    packed.values["qparams"], packed._row_bits, packed.sel, HD // 128, splits=5)
    torch.cuda.synchronize()
    assert proxy.check() >= 15


def test_guard_bands_tcgen05_scoring(guarded_ops):
    rtkv, ops, proxy = guarded_ops
    g = torch.Generator(device="cuda").manual_seed(1)
    for B, H, S, P in ((1, 8, 1000, 128), (2, 3, 130, 26)):
        q = torch.randn(B, H, S, 128, generator=g, device="cuda").to(torch.bfloat16)
        k = torch.randn(B, H, P, 128, generator=g, device="cuda").to(torch.bfloat16)
        lse = torch.full((B, H, S), 8.0, device="cuda")
        ops.prompt_mass_qk(q, k, lse, 128 ** -0.5)
    torch.cuda.synchronize()
    assert proxy.check() >= 4


def test_guard_detects_an_overrun(guarded_ops):
    """The checker itself: one byte written past a guarded buffer is reported."""
    rtkv, ops, proxy = guarded_ops
    t = proxy.empty(3, 5, dtype=torch.float32, device="cuda")
    raw, nbytes = proxy.registry[-1]
    assert proxy.check() == 1
    proxy.empty(3, 5, dtype=torch.float32, device="cuda")
    raw, nbytes = proxy.registry[-1]
    raw[PAD + nbytes + 3] = 0
    with pytest.raises(AssertionError, match="AFTER"):
        proxy.check()
    proxy.registry.clear()


def test_bitwise_repeatability(built_lib):
    """No racecheck on the pool either: every kernel is run 6 times on large inputs (cluster selection kernel, tcgen05
    scoring, group-wise / per-channel / outlier quantisers, packed attention) and must reproduce its first result bit for
    bit; an unsynchronised access would show up as run-to-run differences."""
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import rtkv
    from rtkv import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, S, HD, P = 2, 32, 9000, 4096, 128
    A = torch.softmax(torch.randn(B, H, S, P, generator=g, device="cuda") * 4, dim=-1)
    K = torch.randn(B, S, HD, generator=g, device="cuda")
    V = torch.randn(B, S, HD, generator=g, device="cuda") * 2 + 1
    q = torch.randn(B, H, S, 128, generator=g, device="cuda").to(torch.bfloat16)
    kp = torch.randn(B, H, P, 128, generator=g, device="cuda").to(torch.bfloat16)
    lse = torch.full((B, H, S), 9.0, device="cuda")
    qd = torch.randn(B, H, 128, generator=g, device="cuda")
    cfg = rtkv.CompressionConfig(high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2, later_layer_ratio=0.33)
    comp = rtkv.RealTimePrefillCompressor(cfg)
    ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
    opts = [ops.quant_options("token_group", "token_group", group_size=128, outlier_tau=5.0),
            ops.quant_options("channel", "channel", token_group=32), ops.quant_options("token", "token", outlier_tau=4.0)]

    def run():
        out = []
        k, v, info = comp.compress_layer_kv_cache(K, V, A, ids, 30)          # S >= 8192: cluster selection kernel
        out += [k, v, info["propagation_info"]["selection_mask"].to(torch.uint8)]
        out.append(ops.prompt_mass_qk(q, kp, lse, 128 ** -0.5))
        params, _ = comp.layer_params(30, S, ops.F32)
        sel = ops.score_select(params, B, S, K.device, mass=ops.prompt_mass(A, prompt_len=P))
        ms = int(sel.read_stats()[0]["max_selected"])
        for o in opts:
            out += list(ops.quant_gather_ex(K, V, sel, sel.labels, (2, 4, 8), ms, o))
        pk = ops.quant_pack_ex(K, sel, sel.labels, (2, 4, 8), ms, opts[0], outlier_cap=8)
        out += [pk["codes"], pk["qparams"], pk["out_idx"], pk["out_val"], pk["out_cnt"]]
        packed = comp.compress_layer_kv_cache_packed(K, V, A, ids, 30)
        out.append(packed.attend(qd))
        return [t.clone() for t in out]

    first = run()
    for _ in range(5):
        again = run()
        for a, b in zip(first, again):
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
