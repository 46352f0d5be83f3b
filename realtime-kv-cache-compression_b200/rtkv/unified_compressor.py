"""RealTimePrefillCompressor — drop-in for src/compression/unified_compressor.py.

compress_layer_kv_cache(K, V, attention, input_ids, layer_idx) runs three kernels on
the caller's CUDA stream:
    K1  prompt_mass      attention -> per-token prompt mass
    K2  score_select     mass -> scores, labels, budgeted selection, statistics
    K3  quant_gather     kept K/V rows -> fake-quantised, compacted [B,S',HD]
with ONE host synchronisation in between (S' = max_selected is a tensor shape).
Host tensors go through the C-ABI host entry (rtkv_compress_layer_host) and
come back as host tensors.  There is no CPU path.
"""
import math
import os
import time

import numpy as np
import psutil
import torch

from . import ops
from .dynamic_quantization import DynamicPrecisionQuantizer
from .selective_propagation import SelectiveTokenPropagator, selection_info_from_rows
from .token_importance import LayerWiseImportanceTracker


def _importance_stats(rows, n_tokens):
    """mean / unbiased std / min / max over all B*S scores (unified_compressor.py:158-163)."""
    s = float(rows["score_sum"].sum())
    ss = float(rows["score_sumsq"].sum())
    mean = s / n_tokens
    if n_tokens > 1:
        var = (ss - s * s / n_tokens) / (n_tokens - 1)
        std = float("nan") if math.isnan(var) else math.sqrt(max(0.0, var))      # NaN scores give a NaN std, like torch.std
    else:
        std = float("nan")
    mins, maxs = rows["score_min"], rows["score_max"]
    mn = float("nan") if np.isnan(mins).any() else float(mins.min())
    mx = float("nan") if np.isnan(maxs).any() else float(maxs.max())
    return {"mean_score": mean, "std_score": std, "min_score": mn, "max_score": mx}


def _precision_stats(rows, n_tokens):
    cnt = rows["label_count"].sum(axis=0)
    high, medium = int(cnt[2]), int(cnt[1])
    low = n_tokens - high - medium
    return {"high_count": high, "medium_count": medium, "low_count": low,
            "high_ratio": high / n_tokens, "medium_ratio": medium / n_tokens, "low_ratio": low / n_tokens}


class PackedKV:
    """One layer's kept K/V rows as packed codes: codes [B,S',row_bytes] u8 (bits per row by label), scale
    parameters, optional outlier lists, the kept token positions (sel.sel_idx, ascending) and labels."""

    def __init__(self, keys, values, sel, max_selected, bits, options, dtype, seq_len, width, rows, layer_idx, ratio):
        self.keys, self.values, self.sel, self.max_selected = keys, values, sel, max_selected
        self.bits, self.options, self.dtype, self.seq_len, self.width = bits, options, dtype, seq_len, width
        self.rows, self.layer_idx, self.ratio = rows, layer_idx, ratio

    def nbytes(self):
        """Bytes of the useful part of the packed representation (each kept row at ITS bit width + parameters
        + outlier lists + positions), the number the container would occupy when rows are stored back to back."""
        total = 0
        lab = self.sel.labels.cpu().numpy()
        idx = self.sel.sel_idx.cpu().numpy()
        for b in range(lab.shape[0]):
            c = int(self.rows["n_selected"][b])
            for cls in range(3):
                k = int((lab[b][idx[b, :c]] == cls).sum())
                total += 2 * k * ops.packed_row_bytes(self.width, self.bits[cls])
            total += 4 * c
        for p in (self.keys, self.values):
            total += p["qparams"].numel() * 4
            if p["out_cnt"] is not None:
                total += int(p["out_cnt"].sum()) * 8 + p["out_cnt"].numel() * 4
        return total

    def attend(self, q, num_kv_heads=None, sm_scale=None, want_lse=False):
        """Decode-step attention of q [B,H,128] over this layer's packed rows without expanding them
        (reference-mode packing only: one scale per kept row, no outlier split)."""
        if self.options is not None and (self.options.mode[0] or self.options.mode[1] or self.options.outlier_tau_q8):
            raise NotImplementedError("attend() reads reference-mode packed rows (one scale per kept row)")
        if getattr(self, "_row_bits", None) is None:
            _, sel_labels = ops.gather_scalars(self.sel, self.max_selected)
            lut = torch.tensor(self.bits, dtype=torch.uint8, device=sel_labels.device)
            self._row_bits = lut[sel_labels].contiguous()
        hkv = num_kv_heads if num_kv_heads is not None else self.width // 128
        return ops.packed_decode_attention(q, self.keys["codes"], self.keys["qparams"], self.values["codes"],
                                           self.values["qparams"], self._row_bits, self.sel, hkv, sm_scale,
                                           want_lse=want_lse)

    def attend_prefill(self, q, num_kv_heads=None, sm_scale=None, causal=True, want_lse=False):
        """Prefill attention of q [B,H,Sq,128] (bf16/f16) over this layer's packed rows: the rows are unpacked and dequantised
        once (rtkv_unpack_dequant_ex, any packing mode), rounded to q's dtype, and attended with rtkv_flash_prefill.  Keys are the
        S' kept tokens in ascending position order; `causal` is the top-left aligned rule j <= i over kept-row indices, i.e. the
        reference's attention over K_sel / V_sel with the mask truncated to its first S' columns (modified_llama.py:125-140),
        zero rows of shorter batch rows included, as there.  Returns out [B,H,Sq,128] (and the row LSE)."""
        k, v = self.dequantize()
        B, kept, width = k.shape
        hkv = num_kv_heads if num_kv_heads is not None else width // 128
        k4 = k.to(q.dtype).view(B, kept, hkv, width // hkv).transpose(1, 2)
        v4 = v.to(q.dtype).view(B, kept, hkv, width // hkv).transpose(1, 2)
        out, lse, _ = ops.flash_prefill(q, k4, v4, sm_scale, causal=causal, prompt_len=0, want_lse=want_lse)
        return (out, lse) if want_lse else out

    def dequantize(self):
        k = ops.unpack_dequant_ex(self.keys, self.dtype, self.seq_len, self.width, self.sel, self.sel.labels, self.bits,
                                  self.options, tensor=0)
        v = ops.unpack_dequant_ex(self.values, self.dtype, self.seq_len, self.width, self.sel, self.sel.labels,
                                  self.bits, self.options, tensor=1)
        return k, v


class RealTimePrefillCompressor:
    """reference: unified_compressor.py:12-247"""

    def __init__(self, config, model_config=None):
        self.config = config
        self.model_config = model_config
        self.importance_tracker = LayerWiseImportanceTracker(config)
        self.quantizer = DynamicPrecisionQuantizer(config)
        self.propagator = SelectiveTokenPropagator(config)
        self.compression_stats = {}
        self.layer_states = {}
        self._host_ctx = None
        self._pinned = None

    # -- prompt span ---------------------------------------------------------------
    def identify_prompt_tokens(self, input_ids, special_tokens=None):
        """first max(1, min(S//5, 128)) positions (:35-58)."""
        seq_len = input_ids.shape[1]
        return torch.arange(self.prompt_length(seq_len), device=input_ids.device)

    @staticmethod
    def prompt_length(seq_len):
        return max(1, min(seq_len // 5, 128))

    def extract_attention_weights(self, attention_output, layer_idx):
        """The reference's placeholder fabricates random attention when none is attached
        (:60-93, and its fallback line is broken); here the caller must provide it."""
        if hasattr(attention_output, "attentions"):
            return attention_output.attentions
        raise ValueError("attention weights are not available on this output; run the model with "
                         "output_attentions=True or pass them to compress_layer_kv_cache")

    # -- parameters ------------------------------------------------------------------
    def layer_params(self, layer_idx, seq_len, score_dtype=ops.F32):
        ratio = self.propagator.get_layer_propagation_ratio(layer_idx)
        return ops.make_params(self.config, layer_idx, seq_len, self.prompt_length(seq_len), ratio,
                               score_dtype), ratio

    def _bits(self):
        c = self.config
        return (c.low_precision_bits, c.medium_precision_bits, c.high_precision_bits)

    def _kernel_bits(self, dtype, label_count):
        """Bit widths handed to the quantise kernel.  fp16 rows cannot hold 16-bit codes: the reference's
        torch.clamp(x, 0, 65535) raises on a Half tensor (dynamic_quantization.py:121) - but only for a precision class
        that HAS tokens, empty classes are skipped (:165).  label_count: tokens per class over all B*S tokens."""
        bits = list(self._bits())
        if dtype == torch.float16:
            for c in range(3):
                if bits[c] >= 16:
                    if int(label_count[c]) > 0:
                        raise RuntimeError("value cannot be converted to type at::Half without overflow "
                                           f"({int(label_count[c])} tokens at {bits[c]} bits in float16 rows)")
                    bits[c] = 15          # never used: the class is empty
        return tuple(bits)

    def _quant_options(self):
        """rtkv_quant_options when the config asks for an extended mode, else None (reference semantics)."""
        c = self.config
        if not getattr(c, "extended_quantization", False):
            return None
        return ops.quant_options(c.key_quant_mode, c.value_quant_mode, c.quant_group_size, c.quant_token_group,
                                 c.outlier_tau)

    def _stats_buffer(self, nbytes):
        if self._pinned is None or self._pinned.numel() < nbytes:
            self._pinned = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, pin_memory=True)
        return self._pinned

    # -- the hot path ------------------------------------------------------------------
    def compress_layer_kv_cache(self, key_states, value_states, attention_weights, input_ids, layer_idx,
                                out_keys=None, out_values=None, prompt_mass=None):
        """(:95-172)  key/value [B,S,HD]; attention [B,H,S,>=P]; returns (K', V', info).
        out_keys/out_values (extension, host tensors only): flat buffers of B*S*HD elements the
        result is written into (e.g. pinned memory); the returned tensors are views of them.
        prompt_mass (extension, CUDA): float32 [B,S] prompt-attention mass computed by the caller without a
        materialised attention matrix (ops.prompt_mass_qk from Q, K_prompt and the attention LSE);
        attention_weights may then be None."""
        start_time = time.time()
        B, S, HD = key_states.shape
        if value_states.shape != key_states.shape or value_states.dtype != key_states.dtype:
            raise ValueError("key_states and value_states must have the same shape and dtype")
        P = self.prompt_length(input_ids.shape[1])
        if prompt_mass is not None:
            if not (key_states.is_cuda and prompt_mass.is_cuda) or tuple(prompt_mass.shape) != (B, S):
                raise ValueError("prompt_mass must be a CUDA float32 [batch, seq_len] tensor next to CUDA key/value states")
            params, ratio = self.layer_params(layer_idx, S, ops.F32)
        else:
            if attention_weights.dim() != 4 or attention_weights.shape[0] != B or attention_weights.shape[2] != S:
                raise ValueError("attention_weights must be [batch, heads, seq_len, keys]")
            if P > attention_weights.shape[-1]:
                raise IndexError(f"prompt length {P} exceeds the attention width {attention_weights.shape[-1]}")
            params, ratio = self.layer_params(layer_idx, S, ops.dtype_code(attention_weights))
        params.prompt_len = P
        n_tokens = B * S

        if key_states.is_cuda:
            if not (value_states.is_cuda and (prompt_mass is not None or attention_weights.is_cuda)):
                raise ValueError("key, value and attention tensors must live on the same device")
            dev = key_states.device
            mass = (prompt_mass.float().contiguous() if prompt_mass is not None
                    else ops.prompt_mass(attention_weights, prompt_len=P))
            sel = ops.score_select(params, B, S, dev, mass=mass)
            # one readback (statistics + scores + labels) and one synchronisation: S' is a shape
            nstat = sel.stats.numel()
            off = (nstat + 3) // 4 * 4
            host = self._stats_buffer(off + 5 * n_tokens)
            h_stat = host[:nstat]
            h_scores = host[off: off + 4 * n_tokens].view(torch.float32).view(B, S)
            h_labels = host[off + 4 * n_tokens: off + 5 * n_tokens].view(B, S)
            h_stat.copy_(sel.stats, non_blocking=True)
            h_scores.copy_(sel.scores, non_blocking=True)
            h_labels.copy_(sel.labels, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            raw = h_stat.numpy()
            layer = raw[: ops.LAYER_STATS_BYTES].view(ops._lib.LAYER_STATS)[0]
            rows = raw[ops.LAYER_STATS_BYTES:].view(ops._lib.ROW_STATS).copy()
            max_selected = int(layer["max_selected"])
            qopt = self._quant_options()
            kbits = self._kernel_bits(key_states.dtype, rows["label_count"].sum(axis=0))
            if qopt is None:
                selected_keys, selected_values = ops.quant_gather(
                    key_states, value_states, sel, sel.labels, kbits, max_selected)
            else:
                selected_keys, selected_values = ops.quant_gather_ex(
                    key_states, value_states, sel, sel.labels, kbits, max_selected, qopt)
            scores_cpu = h_scores.clone()
            labels_np = h_labels.numpy().astype(np.int64)
            selection_mask = sel.mask.view(torch.bool)
        else:
            if self._quant_options() is not None:
                raise NotImplementedError("extended quantisation modes need CUDA tensors (rtkv_compress_layer_host "
                                          "implements the reference's per-token scales only)")
            if self._host_ctx is None:
                self._host_ctx = ops.HostContext()
            out = self._host_ctx.compress_layer(key_states, value_states, attention_weights, params,
                                                k_out=out_keys, v_out=out_values)
            rows = out["rows"]
            max_selected = out["max_selected"]
            selected_keys, selected_values = out["k"], out["v"]
            scores_cpu = out["scores"]
            labels_np = out["labels"].numpy().astype(np.int64)
            selection_mask = out["mask"].view(torch.bool)

        self.importance_tracker.layer_scores[layer_idx] = scores_cpu       # token_importance.py:198
        precision_stats = _precision_stats(rows, n_tokens)
        quant_info = {"scales": {}, "zero_points": {}, "bit_assignments": labels_np}
        propagation_info = {
            "layer_idx": layer_idx, "propagation_ratio": ratio, "original_length": S,
            "max_selected_length": max_selected, "selection_mask": selection_mask,
            "selection_stats": selection_info_from_rows(rows, S * ratio)}

        original_memory = key_states.numel() + value_states.numel()
        compressed_memory = selected_keys.numel() + selected_values.numel()
        compression_ratio = compressed_memory / original_memory if original_memory > 0 else 0
        processing_time = time.time() - start_time
        compression_info = {
            "layer_idx": layer_idx,
            "processing_time": processing_time,
            "original_shape": key_states.shape,
            "compressed_shape": selected_keys.shape,
            "compression_ratio": compression_ratio,
            "memory_savings": 1.0 - compression_ratio,
            "importance_stats": _importance_stats(rows, n_tokens),
            "precision_stats": precision_stats,
            "quantization_info": quant_info,
            "propagation_info": propagation_info,
        }
        self.layer_states[layer_idx] = compression_info
        return selected_keys, selected_values, compression_info

    # BASELINE.json wording
    compress = compress_layer_kv_cache

    def compress_layer_kv_cache_packed(self, key_states, value_states, attention_weights, input_ids, layer_idx,
                                       outlier_cap=32):
        """Same selection as compress_layer_kv_cache, but the kept rows are stored as packed integer codes
        (PackedKV) instead of dequantised floats: this is where the memory saving becomes real (the reference
        keeps full-width floats).  PackedKV.dequantize() gives back the tensors
        compress_layer_kv_cache returns (up to the sign of zeros)."""
        B, S, HD = key_states.shape
        P = self.prompt_length(input_ids.shape[1])
        params, ratio = self.layer_params(layer_idx, S, ops.dtype_code(attention_weights))
        params.prompt_len = P
        mass = ops.prompt_mass(attention_weights, prompt_len=P)
        sel = ops.score_select(params, B, S, key_states.device, mass=mass)
        layer, rows = sel.read_stats()
        ms = int(layer["max_selected"])
        qopt = self._quant_options() or ops.quant_options()
        bits = self._kernel_bits(key_states.dtype, rows["label_count"].sum(axis=0))
        pk = ops.quant_pack_ex(key_states, sel, sel.labels, bits, ms, qopt, tensor=0, outlier_cap=outlier_cap)
        pv = ops.quant_pack_ex(value_states, sel, sel.labels, bits, ms, qopt, tensor=1, outlier_cap=outlier_cap)
        return PackedKV(pk, pv, sel, ms, bits, qopt, key_states.dtype, S, HD, rows, layer_idx, ratio)

    def compress_layer_kv_cache_async(self, key_states, value_states, attention_weights, layer_idx,
                                      out_keys, out_values, zero_pad=False, prefetch_stream=None,
                                      prefetch_bytes=40 << 20):
        """Fixed-capacity, synchronisation-free variant for pipelined prefill and CUDA graphs.

        out_keys/out_values: preallocated [B,S,HD]; rows [0, n_selected[b]) of batch row b hold the
        kept tokens, the rest is untouched (or zeroed with zero_pad).  Returns the device-side
        Selection (scores, labels, mask, sel_idx, statistics); nothing is read back.
        prefetch_stream: optional side stream used to prefetch the first prefetch_bytes of K into L2
        while the selection kernel runs."""
        B, S, HD = key_states.shape
        P = self.prompt_length(S)
        params, _ = self.layer_params(layer_idx, S, ops.dtype_code(attention_weights))
        mass = ops.prompt_mass(attention_weights, prompt_len=P)
        if prefetch_stream is not None:
            # while the selection kernel (one SM, latency bound) runs, HBM is idle: warm L2 with the first
            # K rows on a side stream (forked after K1, joined before K3)
            main = torch.cuda.current_stream()
            prefetch_stream.wait_stream(main)
            with torch.cuda.stream(prefetch_stream):
                ops.prefetch_l2(key_states, prefetch_bytes)
        sel = ops.score_select(params, B, S, key_states.device, mass=mass)
        if prefetch_stream is not None:
            torch.cuda.current_stream().wait_stream(prefetch_stream)
        bits = self._bits()
        if key_states.dtype == torch.float16 and max(bits) >= 16:
            # the label counts are not known on the host here: quantise with a representable stand-in width and let
            # collect_async raise (like the reference does) if the 16-bit class turns out to be populated
            bits = tuple(15 if b >= 16 else b for b in bits)
            sel.deferred_fp16_check = True
        ops.quant_gather(key_states, value_states, sel, sel.labels, bits, S, zero_pad=zero_pad,
                         out0=out_keys, out1=out_values)
        return sel

    def collect_async(self, pending):
        """Statistics of a whole prefill in ONE device-to-host transfer.

        pending: list of (layer_idx, Selection, out_keys, out_values) from compress_layer_kv_cache_async.  All
        layers' statistics blocks, scores and labels are concatenated on the device, copied once and the same
        `compression_info` dictionaries as the synchronous call are built from them (and stored in
        `layer_states` / `importance_tracker.layer_scores`).  Returns {layer_idx: (K', V', info)} where K', V'
        are [B, max_selected, HD] views of the fixed-capacity output buffers (rows past a batch row's own
        n_selected are whatever zero_pad left there)."""
        if not pending:
            return {}
        t0 = time.time()
        blobs = []
        for _, sel, _, _ in pending:
            blobs += [sel.stats.view(torch.uint8).reshape(-1), sel.scores.view(torch.uint8).reshape(-1),
                      sel.labels.view(torch.uint8).reshape(-1)]
        dev_blob = torch.cat(blobs)
        host = self._stats_buffer(dev_blob.numel())[: dev_blob.numel()]
        host.copy_(dev_blob, non_blocking=True)
        torch.cuda.current_stream(dev_blob.device).synchronize()
        raw = host.numpy()
        results, off = {}, 0
        share = (time.time() - t0) / len(pending)
        for layer_idx, sel, out_keys, out_values in pending:
            B, S = sel.B, sel.S
            n_tokens = B * S
            nstat = sel.stats.numel()
            st = raw[off: off + nstat]; off += nstat
            scores = torch.from_numpy(raw[off: off + 4 * n_tokens].view(np.float32).reshape(B, S).copy()); off += 4 * n_tokens
            labels_np = raw[off: off + n_tokens].reshape(B, S).astype(np.int64); off += n_tokens
            layer = st[: ops.LAYER_STATS_BYTES].view(ops._lib.LAYER_STATS)[0]
            rows = st[ops.LAYER_STATS_BYTES:].view(ops._lib.ROW_STATS).copy()
            if getattr(sel, "deferred_fp16_check", False):
                self._kernel_bits(torch.float16, rows["label_count"].sum(axis=0))     # raises if the class is populated
            ms = int(layer["max_selected"])
            ratio = self.propagator.get_layer_propagation_ratio(layer_idx)
            k_sel, v_sel = out_keys[:, :ms], out_values[:, :ms]
            original = 2 * out_keys.shape[0] * S * out_keys.shape[2]
            ratio_c = (k_sel.numel() + v_sel.numel()) / original if original else 0
            info = {
                "layer_idx": layer_idx, "processing_time": share,
                "original_shape": torch.Size((B, S, out_keys.shape[2])), "compressed_shape": k_sel.shape,
                "compression_ratio": ratio_c, "memory_savings": 1.0 - ratio_c,
                "importance_stats": _importance_stats(rows, n_tokens),
                "precision_stats": _precision_stats(rows, n_tokens),
                "quantization_info": {"scales": {}, "zero_points": {}, "bit_assignments": labels_np},
                "propagation_info": {"layer_idx": layer_idx, "propagation_ratio": ratio, "original_length": S,
                                     "max_selected_length": ms, "selection_mask": sel.mask.view(torch.bool),
                                     "selection_stats": selection_info_from_rows(rows, S * ratio)},
            }
            self.importance_tracker.layer_scores[layer_idx] = scores
            self.layer_states[layer_idx] = info
            results[layer_idx] = (k_sel, v_sel, info)
        return results

    # -- aggregation (:174-230) ----------------------------------------------------------
    def get_overall_compression_stats(self):
        if not self.layer_states:
            return {}
        states = list(self.layer_states.values())
        n = len(states)
        total_time = sum(s["processing_time"] for s in states)
        high = sum(s["precision_stats"]["high_count"] for s in states)
        medium = sum(s["precision_stats"]["medium_count"] for s in states)
        low = sum(s["precision_stats"]["low_count"] for s in states)
        tokens = high + medium + low
        ordered = sorted(states, key=lambda s: s["layer_idx"])
        cumulative = 1.0
        first_len = ordered[0]["original_shape"][1]
        if first_len > 0:
            cumulative = ordered[-1]["compressed_shape"][1] / first_len
        return {
            "total_layers_processed": n,
            "total_processing_time": total_time,
            "avg_processing_time_per_layer": total_time / n,
            "avg_compression_ratio": sum(s["compression_ratio"] for s in states) / n,
            "avg_memory_savings": sum(s["memory_savings"] for s in states) / n,
            "cumulative_compression": cumulative,
            "overall_memory_savings": 1.0 - cumulative,
            "precision_distribution": {
                "high_ratio": high / tokens if tokens > 0 else 0,
                "medium_ratio": medium / tokens if tokens > 0 else 0,
                "low_ratio": low / tokens if tokens > 0 else 0},
        }

    def reset_compression_state(self):
        """(:232-235)"""
        self.layer_states = {}
        self.importance_tracker.layer_scores = {}

    def estimate_memory_usage(self):
        """(:237-247)"""
        mem = psutil.Process(os.getpid()).memory_info()
        cuda = torch.cuda.is_available()
        return {"rss_mb": mem.rss / (1024 * 1024), "vms_mb": mem.vms / (1024 * 1024),
                "gpu_memory_mb": torch.cuda.memory_allocated() / (1024 * 1024) if cuda else 0,
                "gpu_memory_cached_mb": torch.cuda.memory_reserved() / (1024 * 1024) if cuda else 0}


UnifiedCompressor = RealTimePrefillCompressor


class CompressionHook:
    """Forward-hook registrar kept for API compatibility (reference :250-286 is a print-only
    stub).  The real integration is rtkv.llama_hooks.set_compressor."""

    def __init__(self, compressor):
        self.compressor = compressor
        self.hooks = []

    def register_hooks(self, model):
        for layer_idx, layer in enumerate(model.model.layers):
            self.hooks.append(layer.register_forward_hook(
                lambda module, inputs, output, l=layer_idx: self._compression_hook(module, inputs, output, l)))

    def _compression_hook(self, module, inputs, output, layer_idx):
        return output

    def remove_hooks(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []
