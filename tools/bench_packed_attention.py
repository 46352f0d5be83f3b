"""Decode-step attention over packed rows vs attention over the dequantised fp32 / bf16 rows, at a long-context layer
shape (7B dims, S tokens, the middle-band retention ratio).  Run on the GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import rtkv
from rtkv import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
H, D, HD, PEAK = 32, 128, 4096, 6545.3
gen = torch.Generator(device="cuda").manual_seed(0)
K = torch.randn(1, S, HD, generator=gen, device="cuda")
V = torch.randn(1, S, HD, generator=gen, device="cuda")
P = 128
A = torch.rand(1, H, S, P, generator=gen, device="cuda") / P          # a [..., :P] slice is all the scorer reads
cfg = rtkv.CompressionConfig(num_hidden_layers=32, num_attention_heads=H, hidden_size=HD, low_precision_bits=2,
                             medium_precision_bits=4, high_precision_bits=8)
comp = rtkv.RealTimePrefillCompressor(cfg)
ids = torch.zeros(1, S, dtype=torch.long, device="cuda")
packed = comp.compress_layer_kv_cache_packed(K, V, A, ids, 15)
kept = int(packed.rows["n_selected"][0])
q = torch.randn(1, H, D, generator=gen, device="cuda")
out = packed.attend(q)
bits = packed._row_bits[0, :kept].long()
packed_bytes = int((2 * (HD * bits // 8 + 8) + 1).sum())


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        packed.attend(q)
t_packed = timed(g.replay, 20) / 10
dk, dv = packed.dequantize()
res = {"S": S, "kept_rows": kept, "bits_histogram": {int(b): int((bits == b).sum()) for b in bits.unique()},
       "packed_us": t_packed, "packed_bytes": packed_bytes, "packed_GBs": packed_bytes / t_packed / 1e3,
       "packed_frac_hbm": packed_bytes / t_packed / 1e3 / PEAK}
for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    k4 = dk[:, :kept].view(1, kept, H, D).transpose(1, 2).to(dt).contiguous()
    v4 = dv[:, :kept].view(1, kept, H, D).transpose(1, 2).to(dt).contiguous()
    q4 = q.view(1, H, 1, D).to(dt)
    ref = torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)
    res[f"max_abs_diff_vs_sdpa_{tag}"] = float((ref.float().view(1, H, D) - out).abs().max())
    t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
    res[f"sdpa_{tag}_us"] = t
    res[f"sdpa_{tag}_bytes"] = 2 * kept * HD * k4.element_size()
    res[f"sdpa_{tag}_GBs"] = 2 * kept * HD * k4.element_size() / t / 1e3
print(json.dumps(res))
