"""Budgeted token selection + compacting gather (drop-in for
src/compression/selective_propagation.py)."""
import torch

from . import ops
from ._device import like_input, to_cuda


def selection_info_from_rows(rows, total_budget):
    """selection_stats dict of the reference (:99-159) from the device row statistics.
    Rows that kept nothing contribute nothing to the three lists."""
    info = {"selected_counts": [], "budget_utilization": [], "avg_importance": [],
            "cost_distribution": {"high": 0, "medium": 0, "low": 0}}
    for r in rows:
        n = int(r["n_selected_greedy"])
        if n == 0:
            continue
        info["selected_counts"].append(n)
        info["budget_utilization"].append((int(r["cost_bits"]) / 8.0) / total_budget)
        info["avg_importance"].append(float(r["sel_score_sum"]) / n)
        info["cost_distribution"]["low"] += int(r["sel_class"][0])
        info["cost_distribution"]["medium"] += int(r["sel_class"][1])
        info["cost_distribution"]["high"] += int(r["sel_class"][2])
    if info["selected_counts"]:
        k = len(info["selected_counts"])
        info["avg_selected"] = sum(info["selected_counts"]) / k
        info["avg_budget_util"] = sum(info["budget_utilization"]) / k
        info["overall_avg_importance"] = sum(info["avg_importance"]) / k
    return info


class SelectiveTokenPropagator:
    """reference: selective_propagation.py:6-259"""

    def __init__(self, config):
        self.config = config
        self.early_layer_ratio = config.early_layer_ratio
        self.middle_layer_ratio = config.middle_layer_ratio
        self.later_layer_ratio = config.later_layer_ratio
        total_layers = config.num_hidden_layers
        self.early_boundary = int(0.3 * total_layers)
        self.middle_boundary = int(0.7 * total_layers)

    def get_layer_propagation_ratio(self, layer_idx):
        """depth band -> kept fraction (:23-38)."""
        if layer_idx < self.early_boundary:
            return self.early_layer_ratio
        if layer_idx < self.middle_boundary:
            return self.middle_layer_ratio
        return self.later_layer_ratio

    def _bits(self):
        c = self.config
        return (c.low_precision_bits, c.medium_precision_bits, c.high_precision_bits)

    def compute_token_costs(self, precision_labels):
        """label -> bits/8 as a float tensor (:40-66).  A 3-entry table lookup."""
        lut = torch.tensor([b / 8 for b in self._bits()], dtype=torch.float, device=precision_labels.device)
        return lut[precision_labels.long().clamp(0, 2)]

    def _params(self, seq_len, budget_ratio):
        p = ops.LayerParams()
        p.budget = seq_len * budget_ratio                      # :96
        b = self._bits()
        p.bits[0], p.bits[1], p.bits[2] = int(b[0]), int(b[1]), int(b[2])
        p.fallback_topk = max(1, int(seq_len * 0.1))           # :207
        return p

    def _select(self, importance_scores, precision_labels, budget_ratio, fallback):
        B, S = importance_scores.shape
        sc = to_cuda(importance_scores)
        lab = to_cuda(precision_labels)
        return ops.score_select(self._params(S, budget_ratio), B, S, sc.device, scores=sc, labels=lab,
                                select=True, fallback=fallback)

    def select_tokens_with_budget(self, importance_scores, precision_labels, budget_ratio, layer_idx):
        """greedy walk over descending scores, keep iff cur + cost <= S*ratio (:68-161).
        Ties in the score order break towards the lowest index."""
        B, S = importance_scores.shape
        dev = importance_scores.device
        sel = self._select(importance_scores, precision_labels, budget_ratio, fallback=False)
        _, rows = sel.read_stats()
        mask = sel.mask.view(torch.bool)
        return like_input(mask, dev), selection_info_from_rows(rows, S * budget_ratio)

    def apply_token_selection(self, key_states, value_states, importance_scores, precision_labels,
                              layer_idx, input_ids=None):
        """select, then gather K/V/scores/labels of the kept tokens in ascending index order
        into zero-padded [B, max_selected, .] tensors (:163-244)."""
        ratio = self.get_layer_propagation_ratio(layer_idx)
        B, S, n = key_states.shape
        dev = key_states.device
        sel = self._select(importance_scores, precision_labels, ratio, fallback=True)
        layer, rows = sel.read_stats()
        max_selected = int(layer["max_selected"])
        k, v = to_cuda(key_states), to_cuda(value_states)
        sk, sv = ops.quant_gather(k, v, sel, None, None, max_selected)
        ss, sl = ops.gather_scalars(sel, max_selected)
        info = {"layer_idx": layer_idx, "propagation_ratio": ratio, "original_length": S,
                "max_selected_length": max_selected,
                "selection_mask": like_input(sel.mask.view(torch.bool), dev),
                "selection_stats": selection_info_from_rows(rows, S * ratio)}
        return (like_input(sk, dev), like_input(sv, dev),
                like_input(ss.to(importance_scores.dtype), importance_scores.device),
                like_input(sl.to(precision_labels.dtype), precision_labels.device), info)

    def estimate_compression_ratio(self, layer_idx, original_length):
        """product of the per-layer ratios up to layer_idx (:246-259)."""
        cumulative = 1.0
        for l in range(layer_idx + 1):
            cumulative *= self.get_layer_propagation_ratio(l)
        return {"layer_ratio": self.get_layer_propagation_ratio(layer_idx),
                "cumulative_ratio": cumulative,
                "estimated_length": int(original_length * cumulative),
                "compression_factor": 1.0 / cumulative}
