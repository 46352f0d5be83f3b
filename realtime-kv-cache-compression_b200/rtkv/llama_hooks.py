"""Model-side hook contract of the reference (src/models/modified_llama.py), for the
transformers release in this image (5.x: `position_embeddings`, `past_key_values`, attention
returns `(output, weights)`).

Scope: the *hook* — where the compressor is called, what it is fed,
how its result is used.  The default branch keeps the reference's eager attention formulation in
plain torch around the hook; the fused branch (below) replaces it with librtkv's own attention
kernel.  What is kept from the reference:
  * `set_compressor` chain model -> decoder layer -> attention            (:44-46,183-184,264-268)
  * compression only in eval mode and for q_len > 1 (prefill)             (:102)
  * K/V handed over as [B, S, H_kv*D] contiguous, attention as [B,H,S,S]  (:105-117)
  * on S' != S attention is recomputed on the kept tokens with the additive mask truncated to
    its first S' columns (the reference's positional quirk)               (:125-140)
  * any exception inside the block prints and falls back to the uncompressed result (:144-149)
  * `outputs.compression_info`, `get_compression_stats`, `reset_compression_state`   (:310-339)

Extension (off by default): `set_compressor(c, fused_scoring=True)` replaces the materialised
`softmax(QK^T)` of the prefill by librtkv's own tcgen05 + TMA attention kernel (`ops.flash_prefill`, csrc/flash_prefill.cu):
one launch returns the attention output, the row log-sum-exp and the prompt-attention mass per head (the by-product of the
first key tile that token_importance.py:41-45 takes from the materialised probabilities), so no [B,H,S,S] tensor exists
and the hook works at 32k tokens.  The attention over the kept tokens (:125-142) is the same kernel with kv_len = S'
(top-left aligned causal mask == the reference's mask truncated to its first S' columns).  No library attention kernel is
involved.  Needs bf16/fp16 activations, head_dim 128 and a causal mask (None, or an additive mask that is verified to be
purely causal; any other mask - e.g. padding - takes the eager branch, as does output_attentions=True).
"""
import logging
import math

import torch
from torch import nn

try:  # transformers is only needed for the model wrapper, not for the compressor
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama.modeling_llama import (LlamaAttention, LlamaDecoderLayer,
                                                          apply_rotary_pos_emb, repeat_kv)
    _HAVE_HF = True
except Exception:  # pragma: no cover
    _HAVE_HF = False
    LlamaAttention = LlamaDecoderLayer = LlamaForCausalLM = nn.Module

log = logging.getLogger(__name__)


def _is_causal_mask(mask, q_len, kv_len, cache={}):
    """True if `mask` is None or an additive [B,1,Sq,Skv] mask that is exactly causal (0 on and below the diagonal, a large
    negative number above it).  The same mask object is handed to every layer: the verdict is cached per tensor."""
    if mask is None:
        return True
    key = (mask.data_ptr(), tuple(mask.shape), mask._version)
    if cache.get("key") == key:
        return cache["ok"]
    ok = mask.dim() == 4 and mask.shape[-2] == q_len and mask.shape[-1] >= kv_len and q_len == kv_len
    if ok:
        m = mask[..., :q_len, :kv_len]
        rows = max(1, (1 << 24) // max(1, kv_len))                       # bounded temporaries for long sequences
        j = torch.arange(kv_len, device=mask.device)[None, :]
        for r0 in range(0, q_len, rows):
            r1 = min(q_len, r0 + rows)
            i = torch.arange(r0, r1, device=mask.device)[:, None]
            blk = m[..., r0:r1, :]
            if not (bool((blk.masked_select((j <= i).expand_as(blk)) == 0).all())
                    and bool((blk.masked_select((j > i).expand_as(blk)) <= -1e4).all())):
                ok = False
                break
    cache["key"], cache["ok"] = key, ok
    return ok


def _store_compressed(cache, layer_idx, keys, values):
    """Make the cache entry of this layer the compressed K/V (the reference returns the
    compressed pair as `present_key_value`, :165-166)."""
    layers = getattr(cache, "layers", None)
    if layers is not None and layer_idx < len(layers) and hasattr(layers[layer_idx], "keys"):
        layers[layer_idx].keys = keys
        layers[layer_idx].values = values
        return True
    return False


class CompressedLlamaAttention(LlamaAttention):
    def __init__(self, config, layer_idx=None):
        super().__init__(config, layer_idx)
        self.layer_idx = layer_idx
        self.compressor = None
        self.fused_scoring = False
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.num_key_value_heads = getattr(config, "num_key_value_heads", self.num_heads)

    def set_compressor(self, compressor, fused_scoring=False):
        self.compressor = compressor
        self.fused_scoring = bool(fused_scoring)

    def _forward_fused(self, hidden_states, query_states, key_states, value_states, input_ids):
        """Prefill without the S x S matrix, every attention product on librtkv's tcgen05 kernel:
        flash_prefill (output + LSE + per-head prompt mass) -> compress -> flash_prefill over the kept tokens."""
        from . import ops
        bsz, _, q_len, _ = query_states.shape
        scale = 1.0 / math.sqrt(self.head_dim)
        P = self.compressor.prompt_length(q_len)
        attn_output, _lse, head_sums = ops.flash_prefill(query_states, key_states, value_states, scale, causal=True,
                                                         prompt_len=P, out_layout="bshd")
        try:
            mass = torch.empty(bsz, q_len, dtype=torch.float32, device=hidden_states.device)
            ops.check(ops.lib.rtkv_mass_from_head_sums(ops._ptr(head_sums), bsz, self.num_heads, q_len, ops._ptr(mass),
                                                       ops._stream()))
            k_flat = key_states.transpose(1, 2).contiguous().view(bsz, q_len, -1)
            v_flat = value_states.transpose(1, 2).contiguous().view(bsz, q_len, -1)
            ids = input_ids if input_ids is not None else torch.zeros((bsz, q_len), device=hidden_states.device,
                                                                       dtype=torch.long)
            ck, cv, _info = self.compressor.compress_layer_kv_cache(k_flat, v_flat, None, ids, self.layer_idx,
                                                                    prompt_mass=mass)
            s_kept = ck.shape[1]
            ck = ck.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
            cv = cv.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
            if s_kept != q_len:     # attention over the kept tokens, mask truncated to its first S' columns (:125-140)
                attn_output, _, _ = ops.flash_prefill(query_states, ck, cv, scale, causal=True, want_lse=False,
                                                      out_layout="bshd")
            else:                   # original keys, fake-quantised values (:141-142: attn_weights @ V_quant)
                attn_output, _, _ = ops.flash_prefill(query_states, key_states, cv, scale, causal=True, want_lse=False,
                                                      out_layout="bshd")
            return attn_output, (ck, cv)
        except Exception as e:  # same downgrade as the reference
            print(f"Compression failed at layer {self.layer_idx}: {e}")
            return attn_output, None

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None,
                input_ids=None, **kwargs):
        bsz, q_len, _ = hidden_states.size()
        shape = (bsz, q_len, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(shape).transpose(1, 2)
        if position_embeddings is not None:
            cos, sin = position_embeddings
            query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        if past_key_values is not None:
            key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)

        if (self.fused_scoring and self.compressor is not None and not self.training and q_len > 1
                and hidden_states.is_cuda and query_states.dtype in (torch.bfloat16, torch.float16)
                and self.head_dim == 128 and key_states.shape[2] == q_len and not kwargs.get("output_attentions", False)
                and _is_causal_mask(attention_mask, q_len, key_states.shape[2])):
            attn_output, kept = self._forward_fused(hidden_states, query_states, key_states, value_states, input_ids)
            if kept is not None and past_key_values is not None:
                _store_compressed(past_key_values, self.layer_idx, kept[0], kept[1])
            return self.o_proj(attn_output.reshape(bsz, q_len, -1)), None     # the kernel wrote [B,S,H,D]: no transpose

        groups = self.num_heads // self.num_key_value_heads
        k_full, v_full = repeat_kv(key_states, groups), repeat_kv(value_states, groups)
        attn_weights = torch.matmul(query_states, k_full.transpose(2, 3)) / math.sqrt(self.head_dim)
        if attention_mask is not None:
            attn_weights = attn_weights + attention_mask[..., : k_full.shape[-2]]
        attn_weights = nn.functional.softmax(attn_weights, dim=-1, dtype=torch.float32).to(query_states.dtype)

        if self.compressor is not None and not self.training and q_len > 1:
            try:
                kv_len = key_states.shape[2]
                k_flat = key_states.transpose(1, 2).contiguous().view(bsz, kv_len, -1)
                v_flat = value_states.transpose(1, 2).contiguous().view(bsz, kv_len, -1)
                ids = input_ids if input_ids is not None else torch.zeros(
                    (bsz, q_len), device=hidden_states.device, dtype=torch.long)
                ck, cv, _info = self.compressor.compress_layer_kv_cache(k_flat, v_flat, attn_weights, ids,
                                                                        self.layer_idx)
                s_kept = ck.shape[1]
                ck = ck.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
                cv = cv.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
                if s_kept != kv_len:
                    w = torch.matmul(query_states, repeat_kv(ck, groups).transpose(2, 3)) / math.sqrt(self.head_dim)
                    if attention_mask is not None and s_kept < attention_mask.shape[-1]:
                        w = w + attention_mask[..., :s_kept]
                    w = nn.functional.softmax(w, dim=-1, dtype=torch.float32).to(query_states.dtype)
                    attn_output = torch.matmul(w, repeat_kv(cv, groups))
                else:
                    attn_output = torch.matmul(attn_weights, repeat_kv(cv, groups))
                if past_key_values is not None:
                    _store_compressed(past_key_values, self.layer_idx, ck, cv)
            except Exception as e:  # same downgrade as the reference
                print(f"Compression failed at layer {self.layer_idx}: {e}")
                attn_output = torch.matmul(attn_weights, v_full)
        else:
            attn_output = torch.matmul(attn_weights, v_full)

        attn_output = attn_output.transpose(1, 2).contiguous().reshape(bsz, q_len, -1)
        return self.o_proj(attn_output), attn_weights


class CompressedLlamaDecoderLayer(LlamaDecoderLayer):
    def __init__(self, config, layer_idx):
        super().__init__(config, layer_idx)
        self.self_attn = CompressedLlamaAttention(config, layer_idx)
        self.layer_idx = layer_idx

    def set_compressor(self, compressor, fused_scoring=False):
        self.self_attn.set_compressor(compressor, fused_scoring)


class CompressedLlamaForCausalLM(LlamaForCausalLM):
    def __init__(self, config):
        if getattr(config, "_attn_implementation", None) != "eager":
            config._attn_implementation = "eager"      # the hook needs an additive mask and the weights
        super().__init__(config)
        self.model.layers = nn.ModuleList(
            [CompressedLlamaDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.compressor = None
        self.compression_stats = {}

    def set_compressor(self, compressor, fused_scoring=False):
        self.compressor = compressor
        for layer in self.model.layers:
            layer.set_compressor(compressor, fused_scoring)

    def forward(self, input_ids=None, inputs_embeds=None, **kwargs):
        if input_ids is not None:
            self._current_input_ids = input_ids
        elif inputs_embeds is not None:
            self._current_input_ids = torch.zeros(inputs_embeds.shape[:2], dtype=torch.long,
                                                  device=inputs_embeds.device)
        outputs = super().forward(input_ids=input_ids, inputs_embeds=inputs_embeds, **kwargs)
        try:
            stats = self.compressor.get_overall_compression_stats() if self.compressor is not None else {}
            setattr(outputs, "compression_info", stats)
        except Exception as e:  # never block inference on bookkeeping
            log.warning("Could not attach compression_info: %s", e)
        return outputs

    def get_compression_stats(self):
        return self.compressor.get_overall_compression_stats() if self.compressor is not None else {}

    def reset_compression_state(self):
        if self.compressor is not None:
            self.compressor.reset_compression_state()


def create_compressed_llama_model(model_name, config, device="cuda"):
    """Factory of the reference (:349-386): HF config -> compressed model -> weights (strict=False)
    -> compressor -> device, eval."""
    from .unified_compressor import RealTimePrefillCompressor
    model_config = LlamaConfig.from_pretrained(model_name)
    model = CompressedLlamaForCausalLM(model_config)
    try:
        base = LlamaForCausalLM.from_pretrained(model_name, torch_dtype=torch.float16)
        model.load_state_dict(base.state_dict(), strict=False)
    except Exception as e:
        print(f"Warning: Could not load pre-trained weights: {e}")
    model.set_compressor(RealTimePrefillCompressor(config, model_config))
    model = model.to(device)
    model.eval()
    return model
