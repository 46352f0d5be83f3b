"""ctypes front-end of the CPU oracle (oracle/rtkv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.  Arrays are numpy; bf16/fp16 tensors travel as uint16 bit
patterns together with a dtype code (F32/F16/BF16).
"""
import ctypes
import os
import subprocess

import numpy as np

F32, F16, BF16 = 0, 1, 2
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librtkv_oracle.so")
_lib = None


def build(force=False):
    """Compile rtkv_oracle.c -> librtkv_oracle.so (gcc; OpenMP when available)."""
    src = os.path.join(_HERE, "rtkv_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    base = ["-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + base + omp + ["-o", _SO, src, "-lm"], check=True,
                               capture_output=True, text=True)
                return _SO
            except (OSError, subprocess.CalledProcessError) as e:  # try the next recipe
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_context_relevance.restype = ctypes.c_float
        _lib.orc_compress_layer.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _c(a, dtype=None):
    a = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)
    return a


def _bits3(bits):
    return (ctypes.c_int * 3)(*[int(b) for b in bits])


def np_dtype(dt):
    return np.float32 if dt == F32 else np.uint16


def from_torch(t):
    """torch CPU tensor -> (numpy array, dtype code). Half types become uint16 bit patterns."""
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy(), F32
    if t.dtype == torch.float16:
        return t.view(torch.int16).numpy().view(np.uint16), F16
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16), BF16
    raise TypeError(t.dtype)


def to_torch(a, dt):
    import torch
    if dt == F32:
        return torch.from_numpy(np.ascontiguousarray(a))
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16))
    return t.view(torch.float16 if dt == F16 else torch.bfloat16)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """Override OMP_NUM_THREADS (torchrun pins it to 1) for the CPU-baseline legs of bench.py."""
    lib().orc_set_num_threads(int(n))


def prompt_mass(A, adt, P, idx=None):
    """A: [B,H,S,>=P] (last dim contiguous) -> mass [B,S] float32."""
    assert A.ndim == 4 and A.strides[-1] == A.itemsize
    B, H, S, _ = A.shape
    sb, sh, si = (s // A.itemsize for s in A.strides[:3])
    mass = np.empty((B, S), np.float32)
    if idx is not None:
        idx = _c(idx, np.int64)
        P = len(idx)
    lib().orc_prompt_mass(_p(A), adt, B, H, S, ctypes.c_int64(sb), ctypes.c_int64(sh),
                          ctypes.c_int64(si), _p(idx), int(P), _p(mass))
    return mass


def prompt_mass_groups(A, adt, P, idx=None):
    """fp32 only: (mass [B,S], per-head-group partial sums [B,S,8])."""
    assert A.ndim == 4 and A.strides[-1] == A.itemsize
    B, H, S, _ = A.shape
    sb, sh, si = (s // A.itemsize for s in A.strides[:3])
    mass = np.empty((B, S), np.float32)
    parts = np.zeros((B, S, 8), np.float32)
    if idx is not None:
        idx = _c(idx, np.int64)
        P = len(idx)
    lib().orc_prompt_mass_groups(_p(A), adt, B, H, S, ctypes.c_int64(sb), ctypes.c_int64(sh),
                                 ctypes.c_int64(si), _p(idx), int(P), _p(mass), _p(parts))
    return mass, parts


def prompt_mass_partials(A_local, P, group_size):
    """fp32 local head slice [B,H_local,S,>=P] -> per-group partial sums [B,S,H_local/group_size]."""
    assert A_local.ndim == 4 and A_local.dtype == np.float32 and A_local.strides[-1] == 4
    B, Hl, S, _ = A_local.shape
    sb, sh, si = (s // 4 for s in A_local.strides[:3])
    parts = np.zeros((B, S, Hl // group_size), np.float32)
    lib().orc_prompt_mass_partials(_p(A_local), B, Hl, S, ctypes.c_int64(sb), ctypes.c_int64(sh),
                                   ctypes.c_int64(si), int(P), int(group_size), _p(parts))
    return parts


def mass_from_partials(partials, H):
    partials = _c(partials, np.float32)
    mass = np.empty(partials.shape[:-1], np.float32)
    lib().orc_mass_from_partials(_p(partials), ctypes.c_int64(mass.size), int(partials.shape[-1]), int(H), _p(mass))
    return mass


def prompt_mass_qk(Q, Kp, lse, dt, scale, causal, group_size):
    """Q [B,H,S,D], Kp [B,H,P,D] (storage dtype dt), lse [B,H,S] f32 -> partials [B,S,H/group_size]."""
    Q, Kp, lse = _c(Q), _c(Kp), _c(lse, np.float32)
    B, H, S, D = Q.shape
    P = Kp.shape[2]
    parts = np.zeros((B, S, H // group_size), np.float32)
    lib().orc_prompt_mass_qk(_p(Q), _p(Kp), _p(lse), int(dt), B, H, S, D, P, ctypes.c_double(scale),
                             int(bool(causal)), int(group_size), _p(parts))
    return parts


def position_bias(S):
    pb = np.empty((S,), np.float32)
    lib().orc_position_bias(int(S), _p(pb))
    return pb


def context_relevance(S, P):
    return float(lib().orc_context_relevance(int(S), int(P)))


def normalize(mass, adt=F32):
    mass = _c(mass, np.float32)
    B, S = mass.shape
    out = np.empty_like(mass)
    lib().orc_normalize(_p(mass), adt, B, S, _p(out))
    return out


def importance_scores(mass, P, alpha, beta, gamma, w_l, adt=F32):
    mass = _c(mass, np.float32)
    B, S = mass.shape
    out = np.empty_like(mass)
    lib().orc_importance_scores(_p(mass), adt, B, S, int(P), ctypes.c_double(alpha),
                                ctypes.c_double(beta), ctypes.c_double(gamma),
                                ctypes.c_double(w_l), _p(out))
    return out


def assign_precision(scores, theta_h, theta_m):
    scores = _c(scores, np.float32)
    labels = np.empty(scores.shape, np.uint8)
    counts = np.zeros(3, np.int64)
    lib().orc_assign_precision(_p(scores), ctypes.c_int64(scores.size), ctypes.c_double(theta_h),
                               ctypes.c_double(theta_m), _p(labels), _p(counts))
    return labels, counts


def fake_quant_rows(x, dt, labels, bits, want_codes=False, want_params=False):
    """x: [..., n]; labels: [...] uint8 -> (out, codes|None, qparams|None)."""
    x = _c(x)
    n = x.shape[-1]
    rows = x.size // n
    labels = _c(labels, np.uint8).reshape(-1)
    assert labels.size == rows
    out = np.empty_like(x)
    codes = np.empty((rows, n), np.float32) if want_codes else None
    qp = np.empty((rows, 2), np.float32) if want_params else None
    lib().orc_fake_quant_rows(_p(x), dt, ctypes.c_int64(rows), n, _p(labels), _bits3(bits),
                              _p(out), _p(codes), _p(qp))
    return out, codes, qp


def fake_quant_ex(x, dt, labels, bits, mode=0, group=0, tgroup=0, tau=0.0, cap=-1):
    """Extended granularities / outlier split (project-defined, see orc_fake_quant_ex).
    x: [rows, n] (ONE batch row's kept rows for mode 2); -> dict(out, codes, qparams, omask)."""
    x = _c(x)
    rows, n = x.shape
    labels = _c(labels, np.uint8).reshape(-1)
    assert labels.size == rows
    out = np.empty_like(x)
    codes = np.empty((rows, n), np.float32)
    omask = np.zeros((rows, n), np.uint8)
    if mode == 2:
        qp = np.empty(((rows + tgroup - 1) // tgroup, n, 2), np.float32)
    else:
        g = group if mode == 1 else n
        assert n % g == 0
        qp = np.empty((rows, n // g, 2), np.float32)
    lib().orc_fake_quant_ex(_p(x), dt, ctypes.c_int64(rows), n, _p(labels), _bits3(bits), int(mode), int(group),
                            int(tgroup), int(round(tau * 256)), int(cap), _p(out), _p(codes), _p(qp), _p(omask))
    return dict(out=out, codes=codes, qparams=qp, omask=omask)


def select_tokens(scores, labels, bits, budget):
    scores = _c(scores, np.float32)
    labels = _c(labels, np.uint8)
    B, S = scores.shape
    mask = np.zeros((B, S), np.uint8)
    counts = np.zeros(B, np.int32)
    cost = np.zeros(B, np.float64)
    ssum = np.zeros(B, np.float64)
    cc = np.zeros((B, 3), np.int32)
    lib().orc_select_tokens(_p(scores), _p(labels), B, S, _bits3(bits), ctypes.c_double(budget),
                            _p(mask), _p(counts), _p(cost), _p(ssum), _p(cc))
    return mask, counts, cost, ssum, cc


def topk_mask(scores, k):
    scores = _c(scores, np.float32)
    B, S = scores.shape
    mask = np.zeros((B, S), np.uint8)
    lib().orc_topk_mask(_p(scores), B, S, int(k), _p(mask))
    return mask


def gather_rows(x, dt, mask, max_sel):
    x = _c(x)
    B, S, n = x.shape
    mask = _c(mask, np.uint8)
    out = np.empty((B, max_sel, n), x.dtype)
    lib().orc_gather_rows(_p(x), dt, B, S, n, _p(mask), int(max_sel), _p(out))
    return out


def pack_codes(codes, labels, bits, row_bytes):
    codes = _c(codes, np.float32)
    rows, n = codes.shape
    labels = _c(labels, np.uint8).reshape(-1)
    out = np.zeros((rows, row_bytes), np.uint8)
    lib().orc_pack_codes(_p(codes), ctypes.c_int64(rows), n, _p(labels), _bits3(bits),
                         ctypes.c_int64(row_bytes), _p(out))
    return out


def compress_layer(K, V, dt, A, adt, P, alpha, beta, gamma, w_l, theta_h, theta_m, bits, ratio):
    """Full layer (unified_compressor.py:95-172).  Returns dict of numpy arrays."""
    K = _c(K)
    V = _c(V)
    B, S, HD = K.shape
    assert A.ndim == 4 and A.strides[-1] == A.itemsize
    H = A.shape[1]
    sb, sh, si = (s // A.itemsize for s in A.strides[:3])
    scores = np.empty((B, S), np.float32)
    labels = np.empty((B, S), np.uint8)
    mask = np.empty((B, S), np.uint8)
    counts = np.zeros(B, np.int32)
    wk = np.empty_like(K)
    wv = np.empty_like(V)
    ko = np.empty_like(K)
    vo = np.empty_like(V)
    ms = lib().orc_compress_layer(
        _p(K), _p(V), dt, B, S, HD, _p(A), adt, H, ctypes.c_int64(sb), ctypes.c_int64(sh),
        ctypes.c_int64(si), int(P), ctypes.c_double(alpha), ctypes.c_double(beta),
        ctypes.c_double(gamma), ctypes.c_double(w_l), ctypes.c_double(theta_h),
        ctypes.c_double(theta_m), _bits3(bits), ctypes.c_double(ratio),
        _p(scores), _p(labels), _p(mask), _p(counts), _p(wk), _p(wv), _p(ko), _p(vo))
    k_sel = ko.reshape(-1)[: B * ms * HD].reshape(B, ms, HD).copy()
    v_sel = vo.reshape(-1)[: B * ms * HD].reshape(B, ms, HD).copy()
    return dict(k_sel=k_sel, v_sel=v_sel, scores=scores, labels=labels, mask=mask,
                counts=counts, max_sel=ms, k_quant=wk, v_quant=wv)
