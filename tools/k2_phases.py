#!/usr/bin/env python
"""Per-phase clock64 timestamps of one score_select launch (C2 shape), on the GPU box."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch, numpy as np
import rtkv
from rtkv import ops, _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = 1
dev = torch.device("cuda")
WALK = len(sys.argv) > 2 and sys.argv[2] == "walk"
cfg = (rtkv.CompressionConfig(high_precision_bits=16, medium_precision_bits=8, low_precision_bits=4) if WALK
       else rtkv.CompressionConfig(high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2))
g = torch.Generator(device="cuda").manual_seed(0)
i = torch.arange(S, device="cuda", dtype=torch.float32)
mass = (torch.clamp(128.0 / (i + 1), max=1.0) * (1 + 0.05 * torch.rand(S, generator=g, device="cuda")))[None].contiguous()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
for layer in (0, 12, 31):
    p = ops.make_params(cfg, layer, S, 128, [0.8, 0.6, 0.4][0 if layer < 9 else 1 if layer < 22 else 2], ops.F32)
    for it in range(3):
        _lib.check(_lib.lib.rtkv_debug_k2_timing(ctypes.c_void_p(dbg.data_ptr())))
        sel = ops.score_select(p, B, S, dev, mass=mass)
        torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    n = int(t[0]); ts = t[1:1 + n]
    _, rows = sel.read_stats()
    print(f"layer {layer}: kept {int(rows['n_selected'][0])}/{S} labels {rows['label_count'][0].tolist()} ticks={n} total={int(ts[-1]-ts[0])} cyc; deltas:", np.diff(ts).tolist())
_lib.lib.rtkv_debug_k2_timing(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5): ops.score_select(p, B, S, dev, mass=mass)
e0.record()
for _ in range(50): ops.score_select(p, B, S, dev, mass=mass)
e1.record(); torch.cuda.synchronize()
print("avg K2 launch-to-launch (python loop):", e0.elapsed_time(e1) / 50 * 1e3, "us")
