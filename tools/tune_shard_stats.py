#!/usr/bin/env python
"""Sweep the staged statistics kernel's block size / stage count / bisecting modes (one GPU)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%r, "realtime-kv-cache-compression_b200"))
import torch
import rtkv
from rtkv import ops, dist as rd
dev = torch.device("cuda", 0)
R, variant = int(sys.argv[1]), int(sys.argv[2])
B, H, S, D, P = 1, 32, int(os.environ.get("S", 32768)), 128, 128
hl, w = H // R, H * D // R
g = torch.Generator(device="cuda").manual_seed(1)
As = [torch.rand(B, hl, S, P, generator=g, device=dev) / P for _ in range(3)]
Ks = [torch.randn(B, S, w, generator=g, device=dev) for _ in range(3)]
Vs = [torch.randn(B, S, w, generator=g, device=dev) for _ in range(3)]
v0 = rd.LocalExchange(B, S, R, dev).view(0)
ops.lib.rtkv_debug_shard_variant(variant)
fn = lambda it: ops.shard_stats(As[it %% 3], Ks[it %% 3], Vs[it %% 3], P, H // 8, v0.ptrs, 0, R, v0.state)
for it in range(3): fn(it)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for it in range(6): fn(it)
ts = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3 / 6)
ts.sort()
byts = As[0].numel() * 4 + 2 * Ks[0].numel() * 4
print(json.dumps({"R": R, "variant": variant, "rb": os.environ.get("RTKV_SHARD_RB"), "stages": os.environ.get("RTKV_SHARD_STAGES"),
                  "us": round(ts[len(ts) // 2], 1), "GBs": round(byts / ts[len(ts) // 2] / 1e3)}))
''' % ROOT
for R in (8, 2):
    for S in (1024, 32768):
        env = dict(os.environ, S=str(S))
        r = subprocess.run([sys.executable, "-c", CHILD, str(R), "0"], env=env, capture_output=True, text=True)
        print("S=%d %s" % (S, r.stdout.strip() or r.stderr[-300:]), flush=True)
