"""Model-side hook contract of the reference (src/models/modified_llama.py), for the
transformers release in this image (5.x: `position_embeddings`, `past_key_values`, attention
returns `(output, weights)`).

Scope: the *hook* — where the compressor is called, what it is fed,
how its result is used — not a new attention kernel.  The attention math around the hook
is the reference's eager formulation in plain torch (SURVEY §8f lists fused replacements as
"next").  What is kept from the reference:
  * `set_compressor` chain model -> decoder layer -> attention            (:44-46,183-184,264-268)
  * compression only in eval mode and for q_len > 1 (prefill)             (:102)
  * K/V handed over as [B, S, H_kv*D] contiguous, attention as [B,H,S,S]  (:105-117)
  * on S' != S attention is recomputed on the kept tokens with the additive mask truncated to
    its first S' columns (the reference's positional quirk)               (:125-140)
  * any exception inside the block prints and falls back to the uncompressed result (:144-149)
  * `outputs.compression_info`, `get_compression_stats`, `reset_compression_state`   (:310-339)

Extension (off by default): `set_compressor(c, fused_scoring=True)` replaces the materialised
`softmax(QK^T)` of the prefill by a flash attention that also returns the row log-sum-exp (torch's SDPA flash /
efficient kernels: library code) and takes the prompt-attention mass from librtkv's tcgen05 kernel
(`ops.prompt_mass_qk`: Q, K_prompt, LSE -> mass), so no [B,H,S,S] tensor exists and the hook works at 32k tokens.
Needs bf16/fp16 activations, head_dim 128 and a causal mask; `attn_weights` is then returned as None.
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


def _flash_with_lse(q, k, v, scale):
    """Causal attention output [B,H,S,D] and natural-log row LSE [B,H,S] (float32) from torch's fused SDPA kernels."""
    errors = []
    try:
        r = torch.ops.aten._scaled_dot_product_flash_attention(q, k, v, 0.0, True, False, scale=scale)
        return r[0], r[1].float()
    except Exception as e:  # backend not built for this device / shape
        errors.append(e)
    try:
        r = torch.ops.aten._scaled_dot_product_efficient_attention(q, k, v, None, True, 0.0, True, scale=scale)
        return r[0], r[1][..., : q.shape[2]].float()
    except Exception as e:
        errors.append(e)
    raise RuntimeError(f"no fused attention kernel with LSE output is available: {errors}")


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
        """Prefill without the S x S matrix: flash attention (+LSE) -> tcgen05 prompt mass -> compress ->
        attention over the kept tokens (top-left aligned causal mask == the reference's truncated mask)."""
        from . import ops
        bsz, _, q_len, _ = query_states.shape
        groups = self.num_heads // self.num_key_value_heads
        scale = 1.0 / math.sqrt(self.head_dim)
        k_full, v_full = repeat_kv(key_states, groups), repeat_kv(value_states, groups)
        attn_output, lse = _flash_with_lse(query_states, k_full, v_full, scale)
        try:
            P = self.compressor.prompt_length(q_len)
            mass = ops.prompt_mass_qk(query_states, k_full[:, :, :P], lse, scale, causal=True)
            k_flat = key_states.transpose(1, 2).contiguous().view(bsz, q_len, -1)
            v_flat = value_states.transpose(1, 2).contiguous().view(bsz, q_len, -1)
            ids = input_ids if input_ids is not None else torch.zeros((bsz, q_len), device=hidden_states.device,
                                                                       dtype=torch.long)
            ck, cv, _info = self.compressor.compress_layer_kv_cache(k_flat, v_flat, None, ids, self.layer_idx,
                                                                    prompt_mass=mass)
            s_kept = ck.shape[1]
            ck = ck.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
            cv = cv.view(bsz, s_kept, self.num_key_value_heads, self.head_dim).transpose(1, 2)
            if s_kept != q_len:
                attn_output = nn.functional.scaled_dot_product_attention(
                    query_states, repeat_kv(ck, groups), repeat_kv(cv, groups), is_causal=True, scale=scale)
            else:
                attn_output = nn.functional.scaled_dot_product_attention(
                    query_states, k_full, repeat_kv(cv, groups), is_causal=True, scale=scale)
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
                and self.head_dim == 128 and key_states.shape[2] == q_len):
            attn_output, kept = self._forward_fused(hidden_states, query_states, key_states, value_states, input_ids)
            if kept is not None and past_key_values is not None:
                _store_compressed(past_key_values, self.layer_idx, kept[0], kept[1])
            attn_output = attn_output.transpose(1, 2).contiguous().reshape(bsz, q_len, -1)
            return self.o_proj(attn_output), None

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
