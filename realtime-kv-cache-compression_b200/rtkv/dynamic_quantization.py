"""Dynamic precision assignment + per-token fake quantisation (drop-in for
src/compression/dynamic_quantization.py).  Labels: 0 = LOW, 1 = MID, 2 = HIGH."""
import torch

from . import ops
from ._device import like_input, to_cuda


class DynamicPrecisionQuantizer:
    """reference: dynamic_quantization.py:7-241"""

    def __init__(self, config):
        self.config = config
        self.theta_h = config.theta_h
        self.theta_m = config.theta_m
        self.high_bits = config.high_precision_bits
        self.medium_bits = config.medium_precision_bits
        self.low_bits = config.low_precision_bits

    @property
    def _bits(self):
        return (self.low_bits, self.medium_bits, self.high_bits)

    def _label_counts(self, labels_u8, B, S, device):
        """(low, medium, high) token counts via the row statistics of rtkv_score_select."""
        params = ops.LayerParams()
        params.bits[0] = params.bits[1] = params.bits[2] = 8
        zeros = torch.zeros(B, S, dtype=torch.float32, device=device)
        sel = ops.score_select(params, B, S, device, scores=zeros, labels=labels_u8, select=False)
        _, rows = sel.read_stats()
        c = rows["label_count"].sum(axis=0)
        return int(c[0]), int(c[1]), int(c[2])

    def assign_precision_levels(self, importance_scores):
        """scores [B,S] -> (int64 labels [B,S], stats dict)  (:21-60)."""
        batch_size, seq_len = importance_scores.shape
        dev = importance_scores.device
        sc = to_cuda(importance_scores)
        params = ops.LayerParams()
        params.theta_h, params.theta_m = float(self.theta_h), float(self.theta_m)
        params.bits[0] = params.bits[1] = params.bits[2] = 8
        sel = ops.score_select(params, batch_size, seq_len, sc.device, scores=sc, labels=None, select=False)
        labels = ops.labels_to_i64(sel.labels)
        _, rows = sel.read_stats()
        cnt = rows["label_count"].sum(axis=0)
        total = batch_size * seq_len
        high, medium = int(cnt[2]), int(cnt[1])
        low = total - high - medium
        stats = {"high_count": high, "medium_count": medium, "low_count": low,
                 "high_ratio": high / total, "medium_ratio": medium / total, "low_ratio": low / total}
        return like_input(labels, dev), stats

    def get_quantization_params(self, tensor, num_bits):
        """whole-tensor (scale, zero_point) as 0-dim tensors of tensor.dtype (:62-95)."""
        dev = tensor.device
        qp = ops.quant_params(to_cuda(tensor), num_bits).to(tensor.dtype)
        return like_input(qp[0], dev), like_input(qp[1], dev)

    def quantize_tensor(self, tensor, num_bits, scale, zero_point):
        """clamp(round(x/scale + zp), 0, 2^b-1) dequantised again (:97-126)."""
        dev = tensor.device
        x = to_cuda(tensor)
        qp = torch.stack([torch.as_tensor(scale).reshape(()).to(x.device, torch.float32),
                          torch.as_tensor(zero_point).reshape(()).to(x.device, torch.float32)])
        return like_input(ops.quantize_with_params(x, num_bits, qp), dev)

    def apply_mixed_precision_quantization(self, key_states, value_states, precision_labels):
        """every token's K row and V row quantised at bits[label] (:128-196)."""
        batch_size, seq_len, head_dim = key_states.shape
        dev = key_states.device
        k, v, lab = to_cuda(key_states), to_cuda(value_states), to_cuda(precision_labels)
        qk, qv = ops.quant_gather(k, v, None, lab, self._bits, seq_len)
        quant_info = {"scales": {}, "zero_points": {},
                      "bit_assignments": precision_labels.detach().cpu().numpy()}
        return like_input(qk, dev), like_input(qv, dev), quant_info

    def estimate_memory_savings(self, original_tensor, precision_labels):
        """bytes estimate assuming a 16-bit original (:198-241)."""
        total_elements = original_tensor.numel()
        lab = to_cuda(precision_labels)
        lab2 = lab.reshape(1, -1) if lab.dim() != 2 else lab
        low, medium, high = self._label_counts(lab2.to(torch.uint8), lab2.shape[0], lab2.shape[1], lab2.device)
        width = original_tensor.shape[-1]
        high_e, medium_e, low_e = high * width, medium * width, low * width
        original_memory = total_elements * 2
        compressed_memory = (high_e * (self.high_bits / 8) + medium_e * (self.medium_bits / 8)
                             + low_e * (self.low_bits / 8))
        ratio = compressed_memory / original_memory
        return {"original_memory_mb": original_memory / (1024 * 1024),
                "compressed_memory_mb": compressed_memory / (1024 * 1024),
                "compression_ratio": ratio, "memory_savings": 1.0 - ratio,
                "high_elements_ratio": high_e / total_elements,
                "medium_elements_ratio": medium_e / total_elements,
                "low_elements_ratio": low_e / total_elements}
