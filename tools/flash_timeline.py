"""Timeline of the two softmax warpgroups of rtkv_flash_prefill kernel 4 (CTA 0, warp 0 of each warpgroup): for every own key tile the
cycle at which the tile loop iteration starts, the exponentials start and end.  usage: flash_timeline.py [S] [skew_cycles]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from rtkv import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
skew = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(1, 32, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
buf[62] = 1
ops.lib.rtkv_debug_flash_variant(4)
ops.lib.rtkv_debug_flash_skew(skew)
ops.lib.rtkv_debug_flash_profile(buf.data_ptr())
try:
    for _ in range(2):
        ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
    torch.cuda.synchronize()
finally:
    ops.lib.rtkv_debug_flash_profile(None)
    ops.lib.rtkv_debug_flash_skew(0)
    ops.lib.rtkv_debug_flash_variant(-1)
c = buf.cpu().tolist()
t0 = min(x for x in c[64:64 + 256] if x > 0)
print(f"skew {skew}: own tile | warpgroup 0: start exp_start exp_end (exp len) | warpgroup 1: ... | offset of exp starts")
for j in range(24):
    a = [c[64 + j * 3 + 2] - t0, c[64 + j * 3] - t0, c[64 + j * 3 + 1] - t0]
    b = [c[64 + 128 + j * 3 + 2] - t0, c[64 + 128 + j * 3] - t0, c[64 + 128 + j * 3 + 1] - t0]
    print(f"{j:3d} | {a[0]:7d} {a[1]:7d} {a[2]:7d} ({a[2] - a[1]:5d}) | {b[0]:7d} {b[1]:7d} {b[2]:7d} ({b[2] - b[1]:5d}) | {b[1] - a[1]:6d}")

print("key tile | QK thread: K ready, S buffer free, issued | PV thread: P ready, V ready, issued   (cycles, same origin)")
for t_ in range(4, 20):
    r = [c[384 + t_ * 6 + i] - t0 for i in range(6)]
    print(f"{t_:3d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[3]:7d} {r[4]:7d} {r[5]:7d}")
