"""Tensor-level wrappers over the C ABI (include/rtkv.h).

PyTorch is plumbing here: it owns device memory and the current stream; every
computation is a librtkv kernel.  All functions take CUDA tensors, enqueue on
torch's current stream and do not synchronise unless their docstring says so.
"""
import ctypes
import functools

import numpy as np
import torch

from . import _lib
from ._lib import F32, F16, BF16, FROM_MASS, MAKE_LABELS, SELECT, FALLBACK, LayerParams, check, lib

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}
LAYER_STATS_BYTES = _lib.LAYER_STATS.itemsize
ROW_STATS_BYTES = _lib.ROW_STATS.itemsize


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("rtkv needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def dtype_code(t):
    try:
        return _DT[t.dtype if isinstance(t, torch.Tensor) else t]
    except KeyError:
        raise TypeError(f"rtkv supports float32/float16/bfloat16, got {t.dtype if isinstance(t, torch.Tensor) else t}")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    """The caller's stream on the CURRENT device; every public op runs under _device_guard, which makes the
    device of its tensors current first."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_devices(obj, found):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.append(obj.device)
    elif isinstance(obj, torch.device):
        if obj.type == "cuda":
            found.append(obj if obj.index is not None else torch.device("cuda", torch.cuda.current_device()))
    elif isinstance(obj, Selection):
        found.append(obj.scores.device)
    elif isinstance(obj, dict):
        for v in obj.values():
            _cuda_devices(v, found)


def _device_guard(fn):
    """Launch on the device that owns the tensors, not on whatever device happens to be current: the C entries
    use the current device and torch's current stream of that device (HF device_map / pipeline sharding puts
    layers on several GPUs of one process).  Tensors spread over different devices are refused."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        found = []
        for a in args:
            _cuda_devices(a, found)
        for a in kwargs.values():
            _cuda_devices(a, found)
        if not found:
            return fn(*args, **kwargs)
        dev = found[0]
        for d in found[1:]:
            if d != dev:
                raise ValueError(f"rtkv: tensors live on different devices ({dev} and {d})")
        if dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _check_out(out, like, shape, what):
    """A caller-supplied output buffer must be able to hold the result: same dtype and device as the input,
    contiguous, at least prod(shape) elements."""
    if out is None:
        return
    n = 1
    for d in shape:
        n *= int(d)
    if out.dtype != like.dtype or out.device != like.device or not out.is_contiguous() or out.numel() < n:
        raise ValueError(f"{what}: need a contiguous {like.dtype} buffer on {like.device} with at least {n} elements, got "
                         f"{out.dtype} on {out.device}, {out.numel()} elements, contiguous={out.is_contiguous()}")


_scratch = {}


@_device_guard
def scratch(device, seq_len=1):
    """Per-device scratch (rtkv_scratch_init): header + position-bias log table for seq_len
    entries.  Grown (and re-initialised on the current stream) when a longer sequence shows up."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _scratch.get(key)
    want = max(int(seq_len), 4096)
    if buf is None or (buf.numel() - 256) // 4 < want:
        n = ctypes.c_size_t()
        check(lib.rtkv_scratch_bytes(1, 1 << (want - 1).bit_length(), ctypes.byref(n)))
        buf = torch.empty(n.value, dtype=torch.uint8, device=device)
        check(lib.rtkv_scratch_init(_ptr(buf), buf.numel(), _stream()))
        _scratch[key] = buf
    return buf


def make_params(config, layer_idx, seq_len, prompt_len, ratio, score_dtype=F32):
    """rtkv_layer_params from a CompressionConfig (configs/base_config.py)."""
    p = LayerParams()
    p.alpha, p.beta, p.gamma = float(config.alpha), float(config.beta), float(config.gamma)
    p.layer_weight = float(config.layer_weights[layer_idx])
    p.theta_h, p.theta_m = float(config.theta_h), float(config.theta_m)
    p.budget = seq_len * ratio                      # selective_propagation.py:96
    p.bits[0], p.bits[1], p.bits[2] = (int(config.low_precision_bits), int(config.medium_precision_bits),
                                       int(config.high_precision_bits))
    p.prompt_len = int(prompt_len)
    p.score_dtype = int(score_dtype)
    p.fallback_topk = max(1, int(seq_len * 0.1))    # selective_propagation.py:207
    return p


@_device_guard
def prompt_mass(attn, prompt_len=None, prompt_indices=None):
    """mass[b,i] = sum_p mean_h attn[b,h,i,idx[p]] -> float32 [B,S] (token_importance.py:41-45)."""
    require_cuda()
    assert attn.is_cuda and attn.dim() == 4
    if attn.stride(-1) != 1:
        attn = attn.contiguous()
    B, H, S, W = attn.shape
    idx = None
    if prompt_indices is not None:
        P = int(prompt_indices.numel())
        ar = torch.arange(P, device=prompt_indices.device, dtype=prompt_indices.dtype)
        if not (P <= W and bool(torch.equal(prompt_indices.reshape(-1), ar))):
            idx = prompt_indices.reshape(-1).to(device=attn.device, dtype=torch.int64).contiguous()
            if P and (int(idx.min()) < -W or int(idx.max()) >= W):
                raise IndexError("prompt index out of range")
            idx = torch.where(idx < 0, idx + W, idx)
    else:
        P = int(prompt_len)
        if P > W:
            raise IndexError(f"prompt_len {P} exceeds attention width {W}")
    mass = torch.empty(B, S, dtype=torch.float32, device=attn.device)
    if P == 0:
        return mass.zero_()
    check(lib.rtkv_prompt_mass(_ptr(attn), dtype_code(attn), B, H, S, attn.stride(0), attn.stride(1),
                               attn.stride(2), _ptr(idx), P, _ptr(mass), _stream()))
    return mass


@_device_guard
def prompt_mass_partials(attn_local, prompt_len, group_size):
    """Head-sharded K1a: per-head-group scalar partial sums [B,S,H_local/group_size] (float32 only)."""
    require_cuda()
    if attn_local.stride(-1) != 1:
        attn_local = attn_local.contiguous()
    B, Hl, S, W = attn_local.shape
    if prompt_len > W:
        raise IndexError(f"prompt_len {prompt_len} exceeds attention width {W}")
    parts = torch.empty(B, S, Hl // max(1, group_size), dtype=torch.float32, device=attn_local.device)
    check(lib.rtkv_prompt_mass_partials(_ptr(attn_local), dtype_code(attn_local), B, Hl, S, attn_local.stride(0),
                                        attn_local.stride(1), attn_local.stride(2), None, int(prompt_len),
                                        int(group_size), _ptr(parts), _stream()))
    return parts


@_device_guard
def mass_from_partials(partials, num_heads, rank_major=False):
    """Per-head-group partial sums -> mass = (p0 + p1 + ...)/H with the groups added in head order.
    partials: [B,S,G], or with rank_major=True the all_gather_into_tensor layout [R,B,S,G/R]."""
    require_cuda()
    partials = partials.contiguous()
    assert partials.dtype == torch.float32
    if rank_major:
        R, gpr, shape = partials.shape[0], partials.shape[-1], partials.shape[1:-1]
    else:
        R, gpr, shape = 1, partials.shape[-1], partials.shape[:-1]
    mass = torch.empty(shape, dtype=torch.float32, device=partials.device)
    check(lib.rtkv_mass_from_partials(_ptr(partials), mass.numel(), int(R), int(gpr), int(num_heads), _ptr(mass),
                                      _stream()))
    return mass


@_device_guard
def prompt_mass_qk_partials(q, k_prompt, lse, scale, causal=True):
    """Fused Q.K_prompt^T scoring on tcgen05: q [B,H,S,128], k_prompt [B,H,P,128] (bf16/f16),
    lse [B,H,S] f32 -> per-head prompt-mass sums [B,H,S] (head-major, like lse)."""
    require_cuda()
    q, k_prompt, lse = q.contiguous(), k_prompt.contiguous(), lse.contiguous().float()
    B, H, S, D = q.shape
    P = k_prompt.shape[2]
    parts = torch.empty(B, H, S, dtype=torch.float32, device=q.device)
    check(lib.rtkv_prompt_mass_qk(_ptr(q), _ptr(k_prompt), _ptr(lse), dtype_code(q), B, H, S, D, P,
                                  float(scale), int(bool(causal)), _ptr(parts), _stream()))
    return parts


@_device_guard
def prompt_mass_qk(q, k_prompt, lse, scale, causal=True):
    """mass [B,S] from Q, K_prompt and the row log-sum-exp (no S x S attention matrix)."""
    hs = prompt_mass_qk_partials(q, k_prompt, lse, scale, causal)
    B, H, S = hs.shape
    mass = torch.empty(B, S, dtype=torch.float32, device=hs.device)
    check(lib.rtkv_mass_from_head_sums(_ptr(hs), B, H, S, _ptr(mass), _stream()))
    return mass


@_device_guard
def flash_prefill(q, k, v, sm_scale=None, causal=True, prompt_len=0, want_lse=True, out_layout="bhsd"):
    """Prefill attention on tcgen05 (rtkv_flash_prefill).  q [B,H,Sq,128], k/v [B,Hkv,Skv,128] bf16/f16 views with unit stride
    in the last dim.  Returns (out, lse [B,H,Sq] f32 or None, head_sums [B,H,Sq] f32 or None): head_sums is the per-head
    prompt-column mass (prompt_len > 0), out is [B,H,Sq,128] ("bhsd") or [B,Sq,H,128] ("bshd") contiguous."""
    require_cuda()
    B, H, Sq, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    assert k.shape == v.shape and k.shape[0] == B and k.shape[3] == D and q.dtype == k.dtype == v.dtype
    fix = lambda t: t if (t.stride(3) == 1 and all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0) else t.contiguous()
    q, k, v = fix(q), fix(k), fix(v)
    if sm_scale is None:
        sm_scale = D ** -0.5
    if out_layout == "bshd":
        out = torch.empty(B, Sq, H, D, dtype=q.dtype, device=q.device)
        o_str = (out.stride(0), out.stride(2), out.stride(1))
    else:
        out = torch.empty(B, H, Sq, D, dtype=q.dtype, device=q.device)
        o_str = (out.stride(0), out.stride(1), out.stride(2))
    lse = torch.empty(B, H, Sq, dtype=torch.float32, device=q.device) if want_lse else None
    hs = torch.empty(B, H, Sq, dtype=torch.float32, device=q.device) if prompt_len > 0 else None
    i64x3 = lambda s: (ctypes.c_int64 * 3)(int(s[0]), int(s[1]), int(s[2]))
    check(lib.rtkv_flash_prefill(_ptr(q), _ptr(k), _ptr(v), dtype_code(q), B, H, Hkv, Sq, Skv, D, i64x3(q.stride()),
                                 i64x3(k.stride()), i64x3(v.stride()), float(sm_scale), int(bool(causal)), int(prompt_len),
                                 _ptr(out), i64x3(o_str), _ptr(lse), _ptr(hs), _stream()))
    return out, lse, hs


@_device_guard
def position_bias(seq_len, device):
    require_cuda()
    pb = torch.empty(seq_len, dtype=torch.float32, device=device)
    check(lib.rtkv_position_bias(int(seq_len), _ptr(pb), _stream()))
    return pb


@_device_guard
def normalize_mass(mass, dt=F32):
    require_cuda()
    mass = mass.contiguous()
    out = torch.empty_like(mass)
    B, S = mass.shape
    check(lib.rtkv_normalize_mass(_ptr(mass), int(dt), B, S, _ptr(out), _stream()))
    return out


class Selection:
    """Device-side result of rtkv_score_select."""
    __slots__ = ("scores", "labels", "mask", "sel_idx", "stats", "B", "S", "deferred_fp16_check")

    def __init__(self, scores, labels, mask, sel_idx, stats, B, S):
        self.scores, self.labels, self.mask, self.sel_idx, self.stats = scores, labels, mask, sel_idx, stats
        self.B, self.S = B, S
        self.deferred_fp16_check = False

    @property
    def row_stats_ptr(self):
        return ctypes.c_void_p(self.stats.data_ptr() + LAYER_STATS_BYTES)

    def read_stats(self, pinned=None):
        """D2H of the statistics block + synchronise.  Returns (layer record, row records)."""
        host = pinned if pinned is not None else torch.empty(self.stats.numel(), dtype=torch.uint8, pin_memory=True)
        host = host[: self.stats.numel()]
        host.copy_(self.stats, non_blocking=True)
        torch.cuda.current_stream(self.stats.device).synchronize()
        raw = host.numpy()
        layer = raw[:LAYER_STATS_BYTES].view(_lib.LAYER_STATS)[0]
        rows = raw[LAYER_STATS_BYTES:].view(_lib.ROW_STATS)
        return layer.copy(), rows.copy()


def new_mass_range(batch, device):
    """[B,4] u32 initialised to (0xffffffff, 0, 0, 0): min / max sort key of the mass and NaN flag (rtkv_shard_stats_combine_ranged)."""
    r = torch.zeros(batch, 4, dtype=torch.int32, device=device)
    r[:, 0] = -1
    return r


@_device_guard
def score_select(params, B, S, device, mass=None, scores=None, labels=None, select=True, fallback=True, mass_range=None):
    """Scores -> labels -> budgeted selection (one launch).

    mass given: scores are computed (token_importance.py:134-176); else `scores` is used.
    labels None: labels are computed from the scores (dynamic_quantization.py:38-45)."""
    require_cuda()
    flags = 0
    if mass is not None:
        flags |= FROM_MASS
        mass = mass.contiguous()
        scores = torch.empty(B, S, dtype=torch.float32, device=device)
    else:
        scores = scores.to(torch.float32).contiguous()
    if labels is None:
        flags |= MAKE_LABELS
        labels = torch.empty(B, S, dtype=torch.uint8, device=device)
    else:
        labels = labels.to(torch.uint8).contiguous()
    mask = sel_idx = None
    if select:
        flags |= SELECT
        if fallback:
            flags |= FALLBACK
        mask = torch.empty(B, S, dtype=torch.uint8, device=device)
        sel_idx = torch.empty(B, S, dtype=torch.int32, device=device)
    # the cross-row ticket lives in the layer record and must start at zero when B > 1
    alloc = torch.zeros if B > 1 else torch.empty
    stats = alloc(LAYER_STATS_BYTES + ROW_STATS_BYTES * B, dtype=torch.uint8, device=device)
    scr = scratch(device, S)
    if mass_range is not None:      # the exchange kernel already reduced the row's mass range: K2 skips that pass (and resets it)
        check(lib.rtkv_score_select_ranged(_ptr(mass), _ptr(mass_range), _ptr(scores), _ptr(labels), _ptr(mask), _ptr(sel_idx),
                                           ctypes.c_void_p(stats.data_ptr() + LAYER_STATS_BYTES), _ptr(stats),
                                           _ptr(scr), scr.numel(), B, S, ctypes.byref(params), flags, _stream()))
        return Selection(scores, labels, mask, sel_idx, stats, B, S)
    check(lib.rtkv_score_select(_ptr(mass), _ptr(scores), _ptr(labels), _ptr(mask), _ptr(sel_idx),
                                ctypes.c_void_p(stats.data_ptr() + LAYER_STATS_BYTES), _ptr(stats),
                                _ptr(scr), scr.numel(), B, S, ctypes.byref(params), flags, _stream()))
    return Selection(scores, labels, mask, sel_idx, stats, B, S)


@_device_guard
def quant_gather(x0, x1, sel, labels, bits, out_rows, zero_pad=True, out0=None, out1=None):
    """Fake-quantise (bits given) or copy (bits None) the kept rows of x0/x1 [B,S,n] into
    [B,out_rows,n].  sel: Selection or None (identity: all rows, out_rows <= S)."""
    require_cuda()
    x0 = x0.contiguous()
    B, S, n = x0.shape
    if x1 is not None:
        x1 = x1.contiguous()
        assert x1.shape == x0.shape and x1.dtype == x0.dtype
    _check_out(out0, x0, (B, out_rows, n), "quant_gather out0")
    _check_out(out1, x0, (B, out_rows, n), "quant_gather out1")
    y0 = out0 if out0 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    y1 = None
    if x1 is not None:
        y1 = out1 if out1 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    lab = labels.to(torch.uint8).contiguous() if (labels is not None and bits is not None) else None
    check(lib.rtkv_quant_gather(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n,
                                _ptr(sel.sel_idx) if sel is not None else None,
                                sel.row_stats_ptr if sel is not None else None,
                                _ptr(lab), _lib.bits3(bits) if lab is not None else None,
                                int(out_rows), int(bool(zero_pad)), _ptr(y0), _ptr(y1), _stream()))
    return y0, y1


@_device_guard
def row_minmax(x0, x1, sel, out_rows, negate_max=False):
    """Local min/max of the kept rows: [B,out_rows,4] = (min0, max0, min1, max1) float32
    (max entries negated with negate_max, so that one MIN all-reduce combines ranks)."""
    require_cuda()
    x0, x1 = x0.contiguous(), x1.contiguous()
    B, S, n = x0.shape
    mm = torch.empty(B, out_rows, 4, dtype=torch.float32, device=x0.device)
    check(lib.rtkv_row_minmax(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n,
                              _ptr(sel.sel_idx) if sel is not None else None,
                              sel.row_stats_ptr if sel is not None else None, int(out_rows), int(bool(negate_max)),
                              _ptr(mm), _stream()))
    return mm


@_device_guard
def quant_gather_minmax(x0, x1, sel, labels, bits, out_rows, minmax, zero_pad=True, out0=None, out1=None,
                        negated_max=False):
    """quant_gather with externally supplied (rank-reduced) row min/max."""
    require_cuda()
    x0, x1 = x0.contiguous(), x1.contiguous()
    B, S, n = x0.shape
    _check_out(out0, x0, (B, out_rows, n), "out0")
    _check_out(out1, x0, (B, out_rows, n), "out1")
    y0 = out0 if out0 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    y1 = out1 if out1 is not None else torch.empty_like(y0)
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_quant_gather_minmax(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n,
                                       _ptr(sel.sel_idx) if sel is not None else None,
                                       sel.row_stats_ptr if sel is not None else None, _ptr(lab),
                                       _lib.bits3(bits), int(out_rows), int(bool(zero_pad)),
                                       _ptr(minmax.contiguous()), int(bool(negated_max)), _ptr(y0), _ptr(y1), _stream()))
    return y0, y1


@_device_guard
def quant_gather_p2p(x0, x1, sel, labels, bits, out_rows, mailbox_ptrs, rank, world, state, zero_pad=True,
                     out0=None, out1=None):
    """KV-head-sharded quantise+gather with the row min/max exchanged inside the kernel over
    peer-mapped mailboxes (mailbox_ptrs: list of `world` device addresses, see rtkv.h)."""
    require_cuda()
    x0, x1 = x0.contiguous(), x1.contiguous()
    B, S, n = x0.shape
    _check_out(out0, x0, (B, out_rows, n), "out0")
    _check_out(out1, x0, (B, out_rows, n), "out1")
    y0 = out0 if out0 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    y1 = out1 if out1 is not None else torch.empty_like(y0)
    lab = labels.to(torch.uint8).contiguous()
    ptrs = (ctypes.c_uint64 * world)(*[int(p) for p in mailbox_ptrs])
    check(lib.rtkv_quant_gather_p2p(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n, _ptr(sel.sel_idx), sel.row_stats_ptr,
                                    _ptr(lab), _lib.bits3(bits), int(out_rows), int(bool(zero_pad)), ptrs, int(rank),
                                    int(world), _ptr(state), _ptr(y0), _ptr(y1), _stream()))
    return y0, y1


def shard_exchange_bytes(batch, seq_len, world):
    n = int(lib.rtkv_shard_exchange_bytes(int(batch), int(seq_len), int(world)))
    if n <= 0:
        raise ValueError(f"head sharding needs world_size in (1,2,4,8), got {world}")
    return n


@_device_guard
def shard_stats(attn_local, key_local, value_local, prompt_len, group_size, xbuf_ptrs, rank, world, state):
    """KV-head sharding, producer half of the one exchange per layer: stream the local attention slice
    [B,H/R,S,>=P] (float32) and the local K/V row slices [B,S,n/R] once and push every token's record
    (kmin, kmax, vmin, vmax, prompt-mass partials) into the exchange buffer of every rank (rtkv_shard_stats)."""
    require_cuda()
    if attn_local.stride(-1) != 1:
        attn_local = attn_local.contiguous()
    key_local, value_local = key_local.contiguous(), value_local.contiguous()
    B, Hl, S, W = attn_local.shape
    if prompt_len > W:
        raise IndexError(f"prompt_len {prompt_len} exceeds attention width {W}")
    assert key_local.shape == value_local.shape and key_local.dtype == value_local.dtype and key_local.shape[:2] == (B, S)
    ptrs = (ctypes.c_uint64 * world)(*[int(p) for p in xbuf_ptrs])
    check(lib.rtkv_shard_stats(_ptr(attn_local), dtype_code(attn_local), B, Hl, S, attn_local.stride(0),
                               attn_local.stride(1), attn_local.stride(2), int(prompt_len), int(group_size),
                               _ptr(key_local), _ptr(value_local), dtype_code(key_local), key_local.shape[2],
                               ptrs, int(rank), int(world), _ptr(state), _stream()))


@_device_guard
def shard_stats_combine(attn_local, key_local, value_local, prompt_len, group_size, xbuf_ptrs, rank, world, state, num_heads,
                        mass, minmax, mass_range=None):
    """shard_stats + shard_combine in one kernel (per-CTA peer flags; every rank must run it concurrently):
    fills mass [B,S] and minmax [B,S,4].  Raises RtkvError(RTKV_EUNSUPPORTED) for shapes the staged kernel cannot take."""
    require_cuda()
    if attn_local.stride(-1) != 1:
        attn_local = attn_local.contiguous()
    key_local, value_local = key_local.contiguous(), value_local.contiguous()
    B, Hl, S, W = attn_local.shape
    if prompt_len > W:
        raise IndexError(f"prompt_len {prompt_len} exceeds attention width {W}")
    ptrs = (ctypes.c_uint64 * world)(*[int(p) for p in xbuf_ptrs])
    if mass_range is not None:
        check(lib.rtkv_shard_stats_combine_ranged(_ptr(attn_local), dtype_code(attn_local), B, Hl, S, attn_local.stride(0),
                                                  attn_local.stride(1), attn_local.stride(2), int(prompt_len), int(group_size),
                                                  _ptr(key_local), _ptr(value_local), dtype_code(key_local), key_local.shape[2],
                                                  ptrs, int(rank), int(world), _ptr(state), int(num_heads), _ptr(mass),
                                                  _ptr(minmax), _ptr(mass_range), _stream()))
        return mass, minmax
    check(lib.rtkv_shard_stats_combine(_ptr(attn_local), dtype_code(attn_local), B, Hl, S, attn_local.stride(0),
                                       attn_local.stride(1), attn_local.stride(2), int(prompt_len), int(group_size),
                                       _ptr(key_local), _ptr(value_local), dtype_code(key_local), key_local.shape[2],
                                       ptrs, int(rank), int(world), _ptr(state), int(num_heads), _ptr(mass), _ptr(minmax),
                                       _stream()))
    return mass, minmax


def shard_stats_combine_allranks(attn_slices, key_slices, value_slices, prompt_len, group_size, xbuf_ptrs, states, num_heads,
                                 masses, minmaxes, mass_ranges=None):
    """Test harness (one GPU): rtkv_shard_stats_combine of ALL ranks in one cooperative launch (lists of per-rank tensors)."""
    require_cuda()
    R = len(attn_slices)
    A0 = attn_slices[0]
    B, Hl, S, W = A0.shape
    for t in list(attn_slices) + list(key_slices) + list(value_slices):
        assert t.is_contiguous()
    arr = lambda ts: (ctypes.c_void_p * R)(*[t.data_ptr() for t in ts])
    ptrs = (ctypes.c_uint64 * R)(*[int(p) for p in xbuf_ptrs])
    scratch_p = torch.empty(4096, dtype=torch.uint8, device=A0.device)
    with torch.cuda.device(A0.device):
        check(lib.rtkv_debug_shard_stats_combine_allranks(arr(attn_slices), dtype_code(A0), B, Hl, S, A0.stride(0), A0.stride(1),
                                                          A0.stride(2), int(prompt_len), int(group_size), arr(key_slices),
                                                          arr(value_slices), dtype_code(key_slices[0]), key_slices[0].shape[2],
                                                          ptrs, R, arr(states), int(num_heads), arr(masses), arr(minmaxes),
                                                          arr(mass_ranges) if mass_ranges is not None else None,
                                                          _ptr(scratch_p), _stream()))


@_device_guard
def shard_combine(xbuf_local, state, batch, seq_len, num_heads, world, mass=None, minmax=None):
    """Consumer half: wait for every rank's records, reduce them in rank order -> mass [B,S] f32 and the whole-row
    (kmin, kmax, vmin, vmax) [B,S,4] f32 indexed by token (rtkv_shard_combine)."""
    require_cuda()
    dev = state.device
    if mass is None:
        mass = torch.empty(batch, seq_len, dtype=torch.float32, device=dev)
    if minmax is None:
        minmax = torch.empty(batch, seq_len, 4, dtype=torch.float32, device=dev)
    ptr = xbuf_local if isinstance(xbuf_local, int) else xbuf_local.data_ptr()
    check(lib.rtkv_shard_combine(ctypes.c_void_p(ptr), _ptr(state), int(batch), int(seq_len), int(num_heads), int(world),
                                 _ptr(mass), _ptr(minmax), _stream()))
    return mass, minmax


@_device_guard
def quant_gather_token_minmax(x0, x1, sel, labels, bits, out_rows, minmax_by_token, zero_pad=True, out0=None, out1=None):
    """quant_gather with whole-row min/max given per TOKEN ([B,S,4], from shard_combine)."""
    require_cuda()
    x0, x1 = x0.contiguous(), x1.contiguous()
    B, S, n = x0.shape
    _check_out(out0, x0, (B, out_rows, n), "out0")
    _check_out(out1, x0, (B, out_rows, n), "out1")
    y0 = out0 if out0 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    y1 = out1 if out1 is not None else torch.empty_like(y0)
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_quant_gather_token_minmax(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n,
                                             _ptr(sel.sel_idx) if sel is not None else None,
                                             sel.row_stats_ptr if sel is not None else None, _ptr(lab),
                                             _lib.bits3(bits), int(out_rows), int(bool(zero_pad)),
                                             _ptr(minmax_by_token.contiguous()), _ptr(y0), _ptr(y1), _stream()))
    return y0, y1


@_device_guard
def prefetch_l2(t, nbytes=None):
    """Warm L2 with the first nbytes of tensor t on the current stream (prefetch.global.L2)."""
    require_cuda()
    total = t.numel() * t.element_size()
    n = total if nbytes is None else min(int(nbytes), total)
    check(lib.rtkv_prefetch_l2(_ptr(t), n, _stream()))


@_device_guard
def quant_params(x, bits):
    """Whole-tensor (scale, zero_point) as a float32 [2] device tensor (dynamic_quantization.py:62-95)."""
    require_cuda()
    x = x.contiguous()
    qp = torch.empty(2, dtype=torch.float32, device=x.device)
    words = torch.empty(256, dtype=torch.uint8, device=x.device)   # per call: min / max / NaN flag / ticket words
    check(lib.rtkv_quant_params(_ptr(x), dtype_code(x), x.numel(), int(bits), _ptr(qp), _ptr(words), _stream()))
    return qp


@_device_guard
def quantize_with_params(x, bits, qp):
    require_cuda()
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib.rtkv_quantize_with_params(_ptr(x), dtype_code(x), x.numel(), int(bits), _ptr(qp), _ptr(y), _stream()))
    return y


@_device_guard
def gather_scalars(sel, out_rows):
    """selected_scores [B,out_rows] f32, selected_labels [B,out_rows] i64 (zero padded)."""
    require_cuda()
    dev = sel.scores.device
    ss = torch.empty(sel.B, out_rows, dtype=torch.float32, device=dev)
    sl = torch.empty(sel.B, out_rows, dtype=torch.int64, device=dev)
    check(lib.rtkv_gather_scalars(_ptr(sel.scores), _ptr(sel.labels), _ptr(sel.sel_idx), sel.row_stats_ptr,
                                  sel.B, sel.S, int(out_rows), _ptr(ss), _ptr(sl), _stream()))
    return ss, sl


@_device_guard
def labels_to_i64(labels):
    require_cuda()
    labels = labels.contiguous()
    out = torch.empty(labels.shape, dtype=torch.int64, device=labels.device)
    check(lib.rtkv_labels_to_i64(_ptr(labels), labels.numel(), _ptr(out), _stream()))
    return out


def packed_row_bytes(n, bits):
    v = lib.rtkv_packed_row_bytes(int(n), int(bits))
    if v < 0:
        raise ValueError(f"packed rows need n > 0 and bits in 1..16 (n={n}, bits={bits})")
    return int(v)


@_device_guard
def quant_pack(x, sel, labels, bits, out_rows):
    """Packed integer codes [B,out_rows,row_bytes] u8 + (scale, zero_point) [B,out_rows,2] f32."""
    require_cuda()
    x = x.contiguous()
    B, S, n = x.shape
    rb = packed_row_bytes(n, max(bits))
    codes = torch.empty(B, out_rows, rb, dtype=torch.uint8, device=x.device)
    qp = torch.empty(B, out_rows, 2, dtype=torch.float32, device=x.device)
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_quant_pack(_ptr(x), dtype_code(x), B, S, n, _ptr(sel.sel_idx) if sel is not None else None,
                              sel.row_stats_ptr if sel is not None else None, _ptr(lab), _lib.bits3(bits),
                              int(out_rows), rb, _ptr(codes), _ptr(qp), _stream()))
    return codes, qp


@_device_guard
def unpack_dequant(codes, qp, dtype, S, n, sel, labels, bits):
    require_cuda()
    B, out_rows, rb = codes.shape
    y = torch.empty(B, out_rows, n, dtype=dtype, device=codes.device)
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_unpack_dequant(_ptr(codes), _ptr(qp), dtype_code(dtype), B, S, n,
                                  _ptr(sel.sel_idx) if sel is not None else None,
                                  sel.row_stats_ptr if sel is not None else None, _ptr(lab), _lib.bits3(bits),
                                  int(out_rows), rb, _ptr(y), _stream()))
    return y


_MODES = {"token": _lib.Q_TOKEN, "token_group": _lib.Q_TOKEN_GROUP, "channel": _lib.Q_CHANNEL}


def quant_options(key_mode="token", value_mode="token", group_size=0, token_group=0, outlier_tau=0.0):
    """rtkv_quant_options.  Modes: "token" (the reference: one scale per kept row), "token_group" (per row and
    group of `group_size` channels), "channel" (per channel and group of `token_group` <= 32 kept rows).
    outlier_tau > 0 splits elements with |x| > tau * mean|x| of their group (token modes)."""
    o = _lib.QuantOptions()
    o.mode[0] = _MODES[key_mode] if isinstance(key_mode, str) else int(key_mode)
    o.mode[1] = _MODES[value_mode] if isinstance(value_mode, str) else int(value_mode)
    o.group_size, o.token_group = int(group_size), int(token_group)
    o.outlier_tau_q8 = int(round(float(outlier_tau) * 256))
    return o


@_device_guard
def gather_rows(x, sel, out_rows, zero_pad=True):
    """Coalesced gather of the kept rows of any [B,S,n] tensor (hidden states, K, V) into [B,out_rows,n]
    in ascending token order (selective_propagation.py:213-232)."""
    return quant_gather(x, None, sel, None, None, out_rows, zero_pad)[0]


@_device_guard
def quant_gather_ex(x0, x1, sel, labels, bits, out_rows, options, zero_pad=True, want_counts=False, out0=None,
                    out1=None):
    """quant_gather with extended scale granularities / outlier split (rtkv_quant_gather_ex)."""
    require_cuda()
    x0 = x0.contiguous()
    B, S, n = x0.shape
    if x1 is not None:
        x1 = x1.contiguous()
        assert x1.shape == x0.shape and x1.dtype == x0.dtype
    _check_out(out0, x0, (B, out_rows, n), "quant_gather_ex out0")
    _check_out(out1, x0, (B, out_rows, n), "quant_gather_ex out1")
    y0 = out0 if out0 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    y1 = None
    if x1 is not None:
        y1 = out1 if out1 is not None else torch.empty(B, out_rows, n, dtype=x0.dtype, device=x0.device)
    oc = torch.empty(B, out_rows, 2, dtype=torch.int32, device=x0.device) if want_counts else None
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_quant_gather_ex(_ptr(x0), _ptr(x1), dtype_code(x0), B, S, n,
                                   _ptr(sel.sel_idx) if sel is not None else None,
                                   sel.row_stats_ptr if sel is not None else None, _ptr(lab), _lib.bits3(bits),
                                   int(out_rows), int(bool(zero_pad)), ctypes.byref(options) if options is not None else None,
                                   _ptr(y0), _ptr(y1), _ptr(oc), _stream()))
    return (y0, y1, oc) if want_counts else (y0, y1)


@_device_guard
def quant_pack_ex(x, sel, labels, bits, out_rows, options, tensor=0, outlier_cap=0):
    """Packed codes with options -> dict(codes, qparams, out_idx, out_val, out_cnt, row_bytes)."""
    require_cuda()
    x = x.contiguous()
    B, S, n = x.shape
    rb = packed_row_bytes(n, max(bits))
    mode = options.mode[tensor]
    dev = x.device
    codes = torch.empty(B, out_rows, rb, dtype=torch.uint8, device=dev)
    if mode == _lib.Q_CHANNEL:
        qp = torch.empty(B, (out_rows + options.token_group - 1) // options.token_group, n, 2, dtype=torch.float32, device=dev)
    else:
        qp = torch.empty(B, out_rows, lib.rtkv_quant_groups(n, ctypes.byref(options), tensor), 2, dtype=torch.float32, device=dev)
    split = mode != _lib.Q_CHANNEL and options.outlier_tau_q8 > 0
    cap = int(outlier_cap) if split else 0
    oi = torch.empty(B, out_rows, cap, dtype=torch.int32, device=dev) if split else None
    ov = torch.empty(B, out_rows, cap, dtype=torch.float32, device=dev) if split else None
    ocnt = torch.empty(B, out_rows, dtype=torch.int32, device=dev) if split else None
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_quant_pack_ex(_ptr(x), dtype_code(x), B, S, n, _ptr(sel.sel_idx) if sel is not None else None,
                                 sel.row_stats_ptr if sel is not None else None, _ptr(lab), _lib.bits3(bits),
                                 int(out_rows), rb, ctypes.byref(options), int(tensor), _ptr(codes), _ptr(qp),
                                 _ptr(oi), _ptr(ov), _ptr(ocnt), cap, _stream()))
    return dict(codes=codes, qparams=qp, out_idx=oi, out_val=ov, out_cnt=ocnt, cap=cap)


@_device_guard
def unpack_dequant_ex(packed, dtype, S, n, sel, labels, bits, options, tensor=0):
    require_cuda()
    codes = packed["codes"]
    B, out_rows, rb = codes.shape
    y = torch.empty(B, out_rows, n, dtype=dtype, device=codes.device)
    lab = labels.to(torch.uint8).contiguous()
    check(lib.rtkv_unpack_dequant_ex(_ptr(codes), _ptr(packed["qparams"]), _ptr(packed["out_idx"]),
                                     _ptr(packed["out_val"]), _ptr(packed["out_cnt"]), int(packed["cap"]),
                                     dtype_code(dtype), B, S, n, _ptr(sel.sel_idx) if sel is not None else None,
                                     sel.row_stats_ptr if sel is not None else None, _ptr(lab), _lib.bits3(bits),
                                     int(out_rows), rb, ctypes.byref(options), int(tensor), _ptr(y), _stream()))
    return y


@_device_guard
def packed_decode_attention(q, codes_k, qparams_k, codes_v, qparams_v, row_bits, sel, num_kv_heads, sm_scale=None,
                            splits=None, want_lse=False):
    """Decode-step attention over packed rows (rtkv_packed_decode_attention).  q [B,H,128] -> out [B,H,128] f32."""
    require_cuda()
    q = q.float().contiguous()
    B, H, D = q.shape
    out_rows, rb = codes_k.shape[1], codes_k.shape[2]
    if sm_scale is None:
        sm_scale = D ** -0.5
    if splits is None:                      # ~9 CTAs per SM (3 resident); a CTA takes a multiple of 128 rows (4 warps x 32)
        sms = torch.cuda.get_device_properties(q.device).multi_processor_count
        want = max(1, min((9 * sms + B * H - 1) // (B * H), (out_rows + 127) // 128))
        rows_per_cta = ((out_rows + want - 1) // want + 127) // 128 * 128
        splits = (out_rows + rows_per_cta - 1) // rows_per_cta
    ws = torch.empty(lib.rtkv_packed_attention_workspace_bytes(B, H, splits), dtype=torch.uint8, device=q.device)
    out = torch.empty(B, H, D, dtype=torch.float32, device=q.device)
    lse = torch.empty(B, H, dtype=torch.float32, device=q.device) if want_lse else None
    check(lib.rtkv_packed_decode_attention(_ptr(q), _ptr(codes_k), _ptr(qparams_k), _ptr(codes_v), _ptr(qparams_v),
                                           _ptr(row_bits), sel.row_stats_ptr, rb, B, H, int(num_kv_heads), D, out_rows,
                                           float(sm_scale), int(splits), _ptr(ws), ws.numel(), _ptr(out), _ptr(lse),
                                           _stream()))
    return (out, lse) if want_lse else out


class HostContext:
    """rtkv_ctx: device staging buffers for the host-buffer entry point."""

    def __init__(self):
        require_cuda()
        self._h = ctypes.c_void_p()
        check(lib.rtkv_ctx_create(ctypes.byref(self._h)))
        self._aux = None        # pinned staging for scores / labels / mask (pageable D2H would serialise)

    def close(self):
        if self._h:
            lib.rtkv_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compress_layer(self, K, V, attn, params, k_out=None, v_out=None, want_aux=True):
        """One layer on HOST tensors (synchronous).  Returns dict with k, v ([B,S',n] views of the
        output buffers), max_selected, scores/labels/mask (host) and the statistics records."""
        assert not K.is_cuda and not V.is_cuda and not attn.is_cuda
        K, V = K.contiguous(), V.contiguous()
        if attn.stride(-1) != 1:
            attn = attn.contiguous()
        B, S, n = K.shape
        H = attn.shape[1]
        for name, o in (("k_out", k_out), ("v_out", v_out)):
            if o is not None and (o.is_cuda or o.dtype != K.dtype or not o.is_contiguous() or o.numel() < B * S * n):
                raise ValueError(f"{name}: need a contiguous host {K.dtype} buffer with at least {B * S * n} elements")
        if k_out is None:
            k_out = torch.empty(B * S * n, dtype=K.dtype)
        if v_out is None:
            v_out = torch.empty(B * S * n, dtype=K.dtype)
        scores = labels = mask = None
        if want_aux:
            bs = B * S
            if self._aux is None or self._aux.numel() < 6 * bs:
                self._aux = torch.empty(6 * bs, dtype=torch.uint8, pin_memory=True)
            scores = self._aux[: 4 * bs].view(torch.float32).view(B, S)
            labels = self._aux[4 * bs: 5 * bs].view(B, S)
            mask = self._aux[5 * bs: 6 * bs].view(B, S)
        rows = np.zeros(B, dtype=_lib.ROW_STATS)
        layer = _lib.LayerStats()
        ms = ctypes.c_int32(0)
        check(lib.rtkv_compress_layer_host(
            self._h, _ptr(K), _ptr(V), dtype_code(K), B, S, n, _ptr(attn), dtype_code(attn), H,
            attn.stride(0), attn.stride(1), attn.stride(2), ctypes.byref(params), _ptr(k_out), _ptr(v_out),
            ctypes.byref(ms), _ptr(scores), _ptr(labels), _ptr(mask),
            ctypes.c_void_p(rows.ctypes.data), ctypes.byref(layer)))
        m = int(ms.value)
        if want_aux:            # hand out copies: the pinned staging is reused by the next call
            scores, labels, mask = scores.clone(), labels.clone(), mask.clone()
        return dict(k=k_out.view(-1)[: B * m * n].view(B, m, n), v=v_out.view(-1)[: B * m * n].view(B, m, n),
                    max_selected=m, scores=scores, labels=labels, mask=mask, rows=rows,
                    used_fallback=int(layer.used_fallback))
