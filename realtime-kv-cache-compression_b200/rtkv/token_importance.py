"""Prompt-guided token importance (drop-in for src/compression/token_importance.py).

s_i = alpha * A_hat_i * w_l + beta * log(i+1)/log(T) + gamma * min(1, P/T)

Same class and method names, argument meaning and return types as the reference;
the arithmetic runs in librtkv kernels (prompt_mass.cu, score_select.cu).
"""
import torch

from . import ops
from ._device import like_input, to_cuda


class PromptGuidedImportanceScorer:
    """reference: token_importance.py:7-176"""

    def __init__(self, config):
        self.config = config
        self.alpha = config.alpha
        self.beta = config.beta
        self.gamma = config.gamma
        self.layer_weights = config.layer_weights

    def compute_attention_aggregation(self, attention_weights, prompt_indices, layer_idx):
        """[B,H,S,S'] attention -> [B,S] mass on the prompt key columns (:21-47)."""
        dev = attention_weights.device
        mass = ops.prompt_mass(to_cuda(attention_weights), prompt_indices=prompt_indices)
        return like_input(mass.to(attention_weights.dtype), dev)

    def normalize_attention_scores(self, attention_scores, layer_idx):
        """per-row min-max to [0,1]; 0 where max-min <= 1e-8 (:49-85)."""
        dev, dt = attention_scores.device, attention_scores.dtype
        x = to_cuda(attention_scores)
        out = ops.normalize_mass(x.float(), ops.dtype_code(dt))
        return like_input(out.to(dt), dev)

    def compute_position_bias(self, seq_len, device):
        """log(i+1)/log(T), zeros for T <= 1 (:87-110)."""
        device = torch.device(device)
        ops.require_cuda()
        pb = ops.position_bias(seq_len, device if device.type == "cuda" else torch.device("cuda"))
        return like_input(pb, device) if device.type != "cuda" else pb

    def compute_context_relevance(self, seq_len, prompt_len, device):
        """constant min(1, P/T) (:112-132)."""
        return torch.full((seq_len,), min(1.0, prompt_len / seq_len), device=device)

    def compute_importance_scores(self, attention_weights, prompt_indices, layer_idx):
        """three-term score, float32 [B,S] (:134-176)."""
        B, H, S, _ = attention_weights.shape
        dev = attention_weights.device
        attn = to_cuda(attention_weights)
        P = len(prompt_indices)
        w_l = self.layer_weights[layer_idx]
        mass = ops.prompt_mass(attn, prompt_indices=prompt_indices)
        params = ops.LayerParams()
        params.alpha, params.beta, params.gamma = float(self.alpha), float(self.beta), float(self.gamma)
        params.layer_weight = float(w_l)
        params.prompt_len = int(P)
        params.score_dtype = ops.dtype_code(attention_weights)
        params.bits[0] = params.bits[1] = params.bits[2] = 8
        sel = ops.score_select(params, B, S, attn.device, mass=mass, labels=None, select=False)
        return like_input(sel.scores, dev)


class LayerWiseImportanceTracker:
    """reference: token_importance.py:178-214"""

    def __init__(self, config):
        self.config = config
        self.scorer = PromptGuidedImportanceScorer(config)
        self.layer_scores = {}

    def update_scores(self, layer_idx, attention_weights, prompt_indices):
        scores = self.scorer.compute_importance_scores(attention_weights, prompt_indices, layer_idx)
        self.layer_scores[layer_idx] = scores.detach().cpu()
        return scores

    def get_cumulative_scores(self, layer_idx):
        """mean over the stored layers 0..layer_idx, divided by layer_idx+1 (:202-214).
        Bookkeeping on the host copies the tracker keeps; not part of the hot path."""
        if not self.layer_scores:
            return None
        total = torch.zeros_like(self.layer_scores[0])
        for l in range(min(layer_idx + 1, len(self.layer_scores))):
            if l in self.layer_scores:
                total += self.layer_scores[l]
        return total / (layer_idx + 1)
