#!/usr/bin/env python
"""One launch shape of rtkv_flash_prefill for ncu (small: 8 heads x 8192 tokens)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops
H, S = int(os.environ.get("H", 8)), int(os.environ.get("S", 8192))
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(1, H, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
for _ in range(4):
    ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
torch.cuda.synchronize()
print("ok")
