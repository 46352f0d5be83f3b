"""Launch every extended-quantisation variant a few times, eagerly, at the C2 layer shape (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch

from rtkv import ops

B, S, HD = 1, 4096, 4096
gen = torch.Generator(device="cuda").manual_seed(0)
K = [torch.randn(B, S, HD, generator=gen, device="cuda") for _ in range(3)]
V = [torch.randn(B, S, HD, generator=gen, device="cuda") for _ in range(3)]
ok, ov = torch.empty_like(K[0]), torch.empty_like(K[0])
labels = torch.randint(0, 3, (B, S), generator=gen, device="cuda", dtype=torch.uint8)
variants = [ops.quant_options("token_group", "token_group", group_size=128),
            ops.quant_options("token_group", "token_group", group_size=128, outlier_tau=6.0),
            ops.quant_options("token", "token", outlier_tau=6.0),
            ops.quant_options("channel", "channel", token_group=32),
            ops.quant_options("channel", "channel", token_group=16)]
for rep in range(3):
    ops.quant_gather(K[rep], V[rep], None, labels, (2, 4, 8), S, out0=ok, out1=ov)
    for opt in variants:
        ops.quant_gather_ex(K[rep], V[rep], None, labels, (2, 4, 8), S, opt, out0=ok, out1=ov)
    ops.quant_pack_ex(K[rep], None, labels, (2, 4, 8), S, variants[0])
    ops.quant_pack_ex(K[rep], None, labels, (2, 4, 8), S, variants[3])
torch.cuda.synchronize()
print("ok")
