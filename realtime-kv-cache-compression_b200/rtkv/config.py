"""Hyper-parameters of the compression path.

Field names and defaults are the public API of the reference's
`CompressionConfig` (configs/base_config.py:4-57); anything a reference user
passes by keyword works unchanged.  `layer_weights` defaults to a linear ramp
from 1.0 (first layer) down to 0.5 (last layer).
"""
from dataclasses import dataclass, field
from typing import List, Optional


def linear_layer_weights(num_layers: int, first: float = 1.0, last: float = 0.5) -> List[float]:
    """w_l = first - (first-last) * l/(L-1); same floating-point expression as
    configs/base_config.py:48-51 for the default (1.0, 0.5)."""
    drop = first - last
    return [first - drop * (l / (num_layers - 1)) for l in range(num_layers)]


@dataclass
class CompressionConfig:
    # model
    model_name: str = "meta-llama/Llama-2-7b-hf"
    max_position_embeddings: int = 4096
    num_hidden_layers: int = 32
    hidden_size: int = 4096
    num_attention_heads: int = 32
    # importance score  s = alpha*A_hat*w_l + beta*b_pos + gamma*r
    alpha: float = 0.4
    beta: float = 0.3
    gamma: float = 0.3
    # precision thresholds: s >= theta_h -> high, theta_m <= s < theta_h -> medium, else low
    theta_h: float = 0.7
    theta_m: float = 0.3
    layer_weights: Optional[List[float]] = None
    # fraction of the token budget kept per depth band (first 30% / middle 40% / last 30% of layers)
    early_layer_ratio: float = 0.8
    middle_layer_ratio: float = 0.6
    later_layer_ratio: float = 0.4
    # bit widths per precision class
    high_precision_bits: int = 16
    medium_precision_bits: int = 8
    low_precision_bits: int = 4
    # bookkeeping fields of the reference (not read by the hot path)
    memory_budget_ratio: float = 0.5
    quality_loss_tolerance: float = 0.05
    context_lengths: Optional[List[int]] = None
    batch_sizes: Optional[List[int]] = None
    # extensions (no reference counterpart; the defaults ARE the reference: one scale per token row, no outlier
    # split).  Modes: "token", "token_group" (per row and group of quant_group_size channels), "channel" (per channel
    # and group of quant_token_group <= 32 kept rows).  outlier_tau > 0: |x| > tau * mean|x| of the group is kept exactly.
    key_quant_mode: str = "token"
    value_quant_mode: str = "token"
    quant_group_size: int = 128
    quant_token_group: int = 32
    outlier_tau: float = 0.0

    def __post_init__(self):
        if self.layer_weights is None:
            self.layer_weights = linear_layer_weights(self.num_hidden_layers)
        if self.context_lengths is None:
            self.context_lengths = [4096, 8192, 16384, 32768]
        if self.batch_sizes is None:
            self.batch_sizes = [1, 4, 8]

    @property
    def extended_quantization(self):
        return self.key_quant_mode != "token" or self.value_quant_mode != "token" or self.outlier_tau > 0

    @property
    def bits(self):
        """(low, medium, high) — the order librtkv indexes by label."""
        return (self.low_precision_bits, self.medium_precision_bits, self.high_precision_bits)
