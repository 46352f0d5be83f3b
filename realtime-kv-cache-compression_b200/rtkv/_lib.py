"""ctypes binding of librtkv.so (C ABI: include/rtkv.h).

There is no fallback of any kind: if the shared library has not been built the
import of this module raises, and every compute call needs a CUDA device.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librtkv.so")

F32, F16, BF16 = 0, 1, 2
FROM_MASS, MAKE_LABELS, SELECT, FALLBACK = 1, 2, 4, 8

c_i32, c_i64, c_f32, c_f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double
c_vp = ctypes.c_void_p


class LayerParams(ctypes.Structure):
    """struct rtkv_layer_params"""
    _fields_ = [("alpha", c_f64), ("beta", c_f64), ("gamma", c_f64), ("layer_weight", c_f64),
                ("theta_h", c_f64), ("theta_m", c_f64), ("budget", c_f64),
                ("bits", c_i32 * 3), ("prompt_len", c_i32), ("score_dtype", c_i32),
                ("fallback_topk", c_i32)]


class QuantOptions(ctypes.Structure):
    """rtkv_quant_options (include/rtkv.h): scale granularity per tensor + outlier split."""
    _fields_ = [("mode", ctypes.c_int32 * 2), ("group_size", ctypes.c_int32), ("token_group", ctypes.c_int32),
                ("outlier_tau_q8", ctypes.c_int32), ("reserved", ctypes.c_int32 * 3)]


Q_TOKEN, Q_TOKEN_GROUP, Q_CHANNEL = 0, 1, 2


class LayerStats(ctypes.Structure):
    """struct rtkv_layer_stats"""
    _fields_ = [("max_selected", c_i32), ("used_fallback", c_i32), ("reserved", c_i32 * 2)]


# struct rtkv_row_stats as a numpy record (72 bytes, natural alignment == packed)
ROW_STATS = np.dtype([
    ("n_selected", "<i4"), ("n_selected_greedy", "<i4"), ("sel_class", "<i4", (3,)),
    ("label_count", "<i4", (3,)), ("cost_bits", "<i8"), ("sel_score_sum", "<f8"),
    ("score_sum", "<f8"), ("score_sumsq", "<f8"), ("score_min", "<f4"), ("score_max", "<f4")])
LAYER_STATS = np.dtype([("max_selected", "<i4"), ("used_fallback", "<i4"), ("reserved", "<i4", (2,))])
assert ROW_STATS.itemsize == 72 and LAYER_STATS.itemsize == 16 and ctypes.sizeof(LayerParams) == 80

# name -> (restype, argtypes); one entry per symbol declared in include/rtkv.h
PROTOTYPES = {
    "rtkv_abi_version": (ctypes.c_int, []),
    "rtkv_last_error": (ctypes.c_char_p, []),
    "rtkv_device_info": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "rtkv_scratch_bytes": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "rtkv_prompt_mass": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        c_i64, c_i64, c_i64, c_vp, ctypes.c_int, c_vp, c_vp]),
    "rtkv_prompt_mass_partials": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 c_i64, c_i64, c_i64, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_mass_from_partials": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_scratch_init": (ctypes.c_int, [c_vp, ctypes.c_size_t, c_vp]),
    "rtkv_prompt_mass_qk": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32, ctypes.c_int,
                                           c_vp, c_vp]),
    "rtkv_mass_from_head_sums": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_position_bias": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp]),
    "rtkv_normalize_mass": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_score_select": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t,
                                         ctypes.c_int, ctypes.c_int, ctypes.POINTER(LayerParams), ctypes.c_int,
                                         c_vp]),
    "rtkv_quant_gather": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                         c_vp, c_vp, c_vp]),
    "rtkv_quant_gather_range": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                               c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               c_vp]),
    "rtkv_row_minmax": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                       ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_quant_gather_minmax": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                                c_vp, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rtkv_p2p_mailbox_bytes": (c_i64, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rtkv_quant_gather_p2p": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, c_vp,
                                             c_vp, c_vp, c_vp]),
    "rtkv_shard_exchange_bytes": (c_i64, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rtkv_shard_stats": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64, c_i64, c_i64,
                                        ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rtkv_shard_stats_combine": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64, c_i64, c_i64,
                                                ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int,
                                                c_vp, c_vp, c_vp]),
    "rtkv_shard_stats_combine_ranged": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64, c_i64, c_i64,
                                                       ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                                       ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int,
                                                       c_vp, c_vp, c_vp, c_vp]),
    "rtkv_score_select_ranged": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t,
                                                ctypes.c_int, ctypes.c_int, ctypes.POINTER(LayerParams), ctypes.c_int, c_vp]),
    "rtkv_shard_combine": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                          c_vp]),
    "rtkv_quant_gather_token_minmax": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                                      c_vp, c_vp, c_vp, c_vp]),
    "rtkv_prefetch_l2": (ctypes.c_int, [c_vp, ctypes.c_size_t, c_vp]),
    "rtkv_quant_params": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rtkv_quantize_with_params": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rtkv_gather_scalars": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           c_vp, c_vp, c_vp]),
    "rtkv_labels_to_i64": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "rtkv_packed_row_bytes": (c_i64, [ctypes.c_int, ctypes.c_int]),
    "rtkv_quant_pack": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                       c_vp, ctypes.POINTER(c_i32), ctypes.c_int, c_i64, c_vp, c_vp, c_vp]),
    "rtkv_unpack_dequant": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, c_i64, c_vp,
                                           c_vp]),
    "rtkv_quant_gather_ex": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(QuantOptions), c_vp, c_vp, c_vp, c_vp]),
    "rtkv_quant_groups": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(QuantOptions), ctypes.c_int]),
    "rtkv_quant_pack_ex": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                          c_vp, ctypes.POINTER(c_i32), ctypes.c_int, c_i64,
                                          ctypes.POINTER(QuantOptions), ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                          ctypes.c_int, c_vp]),
    "rtkv_unpack_dequant_ex": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.POINTER(c_i32),
                                              ctypes.c_int, c_i64, ctypes.POINTER(QuantOptions), ctypes.c_int, c_vp,
                                              c_vp]),
    "rtkv_packed_attention_workspace_bytes": (c_i64, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rtkv_packed_decode_attention": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                    ctypes.c_int, c_vp, ctypes.c_size_t, c_vp, c_vp, c_vp]),
    "rtkv_flash_prefill": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                          ctypes.POINTER(c_i64), ctypes.c_float, ctypes.c_int, ctypes.c_int, c_vp,
                                          ctypes.POINTER(c_i64), c_vp, c_vp, c_vp]),
    "rtkv_debug_channel_variant": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_k2_timing": (ctypes.c_int, [c_vp]),
    "rtkv_debug_k2_mode": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_pdl": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_flash_variant": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_flash_poly": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_flash_profile": (ctypes.c_int, [ctypes.c_void_p]),
    "rtkv_debug_flash_skew": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_mma_rate": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "rtkv_debug_bf16_packed": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_shard_variant": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_shard_grid": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_debug_shard_stats_combine_allranks": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                               ctypes.c_int, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int,
                                                               ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int,
                                                               ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.POINTER(c_vp),
                                                               ctypes.c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_vp,
                                                               c_vp]),
    "rtkv_debug_qk_mode": (ctypes.c_int, [ctypes.c_int]),
    "rtkv_ctx_create": (ctypes.c_int, [ctypes.POINTER(c_vp)]),
    "rtkv_ctx_destroy": (ctypes.c_int, [c_vp]),
    "rtkv_compress_layer_host": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int, c_i64, c_i64,
                                                c_i64, ctypes.POINTER(LayerParams), c_vp, c_vp,
                                                ctypes.POINTER(c_i32), c_vp, c_vp, c_vp, c_vp, c_vp]),
}


class RtkvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"librtkv error {code}: {message}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python realtime-kv-cache-compression_b200/build.py` "
            "(or __graft_entry__.build()).  rtkv has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.rtkv_abi_version() != 1:
        raise ImportError("librtkv.so ABI version mismatch")
    return lib


lib = _load()


def check(code):
    if code != 0:
        raise RtkvError(code, lib.rtkv_last_error().decode("utf-8", "replace"))


def bits3(bits):
    return (c_i32 * 3)(int(bits[0]), int(bits[1]), int(bits[2]))
