"""rtkv_flash_prefill against torch SDPA (cuDNN / flash, library) on causal bf16 attention, 32 heads x head_dim 128.
Prints one JSON line per sequence length: TFLOP/s (causal FLOPs = 4 B H S^2 D / 2) of the automatic kernel choice, of every kernel
variant (rtkv_debug_flash_variant), optionally of the polynomial-exponential shares (--poly) and of SDPA."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops

VARIANTS = {0: "both_warpgroups_per_tile", 1: "independent_key_halves", 2: "two_query_tiles_P_in_TMEM", 3: "alternating_tiles_two_accumulators",
            4: "alternating_tiles_one_accumulator_P_in_TMEM"}


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


lengths = [int(x) for x in sys.argv[1:] if x.isdigit()] or [1024, 2048, 4096, 8192, 32768]
for S in lengths:
    B, H = 1, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, H, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
    flops = 4.0 * B * H * S * S * 128 / 2
    call = lambda: ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
    tf = lambda ms: round(flops / ms / 1e9, 1)
    out = {"B": B, "H": H, "S": S}
    t_auto = timeit(call)
    out["rtkv_ms"], out["rtkv_tflops"] = round(t_auto, 3), tf(t_auto)
    for var, name in VARIANTS.items():
        ops.lib.rtkv_debug_flash_variant(var)
        out[name + "_tflops"] = tf(timeit(call))
        if "--poly" in sys.argv and var in (0, 4):
            for pe in (2, 3, 4):
                ops.lib.rtkv_debug_flash_poly(pe)
                out[f"{name}_poly_every_{pe}_tflops"] = tf(timeit(call))
            ops.lib.rtkv_debug_flash_poly(0)
    ops.lib.rtkv_debug_flash_variant(-1)
    t_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True))
    out["sdpa_ms"], out["sdpa_tflops"] = round(t_sdpa, 3), tf(t_sdpa)
    print(json.dumps(out), flush=True)
