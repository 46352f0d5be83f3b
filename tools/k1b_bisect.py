"""K1b (rtkv_prompt_mass_qk) at the 32k shape with parts of the kernel switched off (rtkv_debug_qk_mode): what bounds it?
mode 0 full kernel; 1 no exponentials (raw logits summed); 2 no TMEM loads / epilogue math at all; 3 no MMAs (TMA + epilogue only)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from rtkv import ops

B, H, S, D, P = 1, 32, 32768, 128, 128
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B, H, S, D, generator=g, device="cuda").to(torch.bfloat16)
kp = torch.randn(B, H, P, D, generator=g, device="cuda").to(torch.bfloat16)
lse = torch.full((B, H, S), 12.0, device="cuda")
out = torch.empty(B, H, S, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    ops.check(ops.lib.rtkv_prompt_mass_qk(q.data_ptr(), kp.data_ptr(), lse.data_ptr(), 2, B, H, S, D, P, D ** -0.5, 1, out.data_ptr(), None))


res = {}
for mode in (0, 1, 2, 3, 0):
    ops.lib.rtkv_debug_qk_mode(mode)
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    res[f"mode{mode}" + ("_again" if f"mode{mode}" in res else "")] = round(sorted(ts)[len(ts) // 2], 1)
ops.lib.rtkv_debug_qk_mode(0)
print(json.dumps({"shape": "c3", "us": res, "hbm_floor_us": round((q.numel() * 2 + lse.numel() * 4 + out.numel() * 4) / 6545.3e3, 1)}))
