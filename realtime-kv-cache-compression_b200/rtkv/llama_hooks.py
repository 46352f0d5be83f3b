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
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.num_key_value_heads = getattr(config, "num_key_value_heads", self.num_heads)

    def set_compressor(self, compressor):
        self.compressor = compressor

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

    def set_compressor(self, compressor):
        self.self_attn.set_compressor(compressor)


class CompressedLlamaForCausalLM(LlamaForCausalLM):
    def __init__(self, config):
        if getattr(config, "_attn_implementation", None) != "eager":
            config._attn_implementation = "eager"      # the hook needs an additive mask and the weights
        super().__init__(config)
        self.model.layers = nn.ModuleList(
            [CompressedLlamaDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.compressor = None
        self.compression_stats = {}

    def set_compressor(self, compressor):
        self.compressor = compressor
        for layer in self.model.layers:
            layer.set_compressor(compressor)

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
