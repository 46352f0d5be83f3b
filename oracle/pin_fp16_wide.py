#!/usr/bin/env python
"""fp16 K/V with the reference's DEFAULT widths 16/8/4 (build container only; imports /root/reference).

torch.clamp(x, 0, 65535) raises on a Half tensor (dynamic_quantization.py:121) - but the reference only reaches it
for a precision class that has tokens (empty classes are skipped, :165).  Two fixtures pin both behaviours:
  layer_f16_16_8_4_nohigh       deep layer, no HIGH token: the reference compresses the layer normally
  layer_f16_16_8_4_high_raises  layer 0 with a late token that takes the whole prompt mass (-> HIGH): the reference raises
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import oracle  # noqa: E402
from oracle.pin_against_reference import GOLD, TDT, causal_probs, golden_layer, run_reference_layer, t2n  # noqa: E402
from configs.base_config import CompressionConfig  # noqa: E402


def main():
    m = golden_layer("layer_f16_16_8_4_nohigh", 8, 1, 8, 128, 64, oracle.F16, (4, 8, 16), 32, 30)
    assert m["precision_stats"]["high_count"] == 0, m["precision_stats"]
    # ---- populated HIGH class: the reference raises
    gen = torch.Generator().manual_seed(9)
    B, H, S, HD, L, layer = 1, 8, 128, 64, 32, 0
    A = causal_probs(B, H, S, torch.float16, gen)
    A[:, :, S - 2, :] = 0.05                     # a late token with by far the largest prompt mass -> n = 1, score >= 0.7
    K = torch.randn(B, S, HD, generator=gen).to(torch.float16)
    V = torch.randn(B, S, HD, generator=gen).to(torch.float16)
    cfg = CompressionConfig(num_hidden_layers=L, num_attention_heads=H, hidden_size=HD)      # defaults: 16/8/4
    try:
        run_reference_layer(cfg, K, V, A, layer)
        raise SystemExit("the reference did not raise")
    except RuntimeError as e:
        msg = str(e)
    assert "Half" in msg, msg
    P = max(1, min(S // 5, 128))
    meta = dict(name="layer_f16_16_8_4_high_raises", B=B, H=H, S=S, HD=HD, P=P, dt=oracle.F16, bits=[4, 8, 16], L=L,
                layer_idx=layer, error=msg)
    np.savez_compressed(os.path.join(GOLD, "layer_f16_16_8_4_high_raises.npz"), K=t2n(K), V=t2n(V),
                        A=t2n(A[..., :P].contiguous()), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print("ok:", m["precision_stats"], "| raises:", msg)


if __name__ == "__main__":
    main()
