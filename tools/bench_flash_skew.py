"""rtkv_flash_prefill kernel 4 at 8k / 32k tokens for a list of warpgroup skews (rtkv_debug_flash_skew).  usage: bench_flash_skew.py 0 1500 2800 ..."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from rtkv import ops


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


skews = [int(x) for x in sys.argv[1:]] or [0, 2800]
ops.lib.rtkv_debug_flash_variant(4)
for S in (8192, 32768):
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(1, 32, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
    flops = 4.0 * 32 * S * S * 128 / 2
    out = {"S": S}
    for sk in skews + skews[:1]:
        ops.lib.rtkv_debug_flash_skew(sk)
        out[f"skew_{sk}" + ("_again" if f"skew_{sk}" in out else "")] = round(flops / timeit(lambda: ops.flash_prefill(q, k, v, causal=True, prompt_len=128)) / 1e9, 1)
    print(json.dumps(out), flush=True)
ops.lib.rtkv_debug_flash_skew(0)
ops.lib.rtkv_debug_flash_variant(-1)
