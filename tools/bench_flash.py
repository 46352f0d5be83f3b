#!/usr/bin/env python
"""rtkv_flash_prefill vs torch SDPA (library flash kernel) on one GPU: causal prefill, bf16, head_dim 128."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for (B, H, S) in [(1, 32, 4096), (1, 32, 8192), (1, 32, 32768)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, H, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
    flops = 4.0 * B * H * S * S * 128 / 2
    t_rtkv = timeit(lambda: ops.flash_prefill(q, k, v, causal=True, prompt_len=128))
    tp = {}
    for pe in (0, 2, 3, 4):
        ops.lib.rtkv_debug_flash_poly(pe)
        tp[pe] = round(flops / timeit(lambda: ops.flash_prefill(q, k, v, causal=True, prompt_len=128)) / 1e9, 1)
    ops.lib.rtkv_debug_flash_poly(0)
    ts = {}
    for var in (1, 2):
        ops.lib.rtkv_debug_flash_variant(var)
        ts[var] = timeit(lambda: ops.flash_prefill(q, k, v, causal=True, prompt_len=128))
    ops.lib.rtkv_debug_flash_variant(0)
    t_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True))
    print(json.dumps({"B": B, "H": H, "S": S, "rtkv_ms": round(t_rtkv, 3), "rtkv_tflops": round(flops / t_rtkv / 1e9, 1), "poly_every_tflops": tp, "independent_halves_variant_tflops": round(flops / ts[1] / 1e9, 1), "two_tile_tmem_p_variant_tflops": round(flops / ts[2] / 1e9, 1),
                      "sdpa_ms": round(t_sdpa, 3), "sdpa_tflops": round(flops / t_sdpa / 1e9, 1)}), flush=True)
