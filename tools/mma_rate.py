"""Cycles per 128x128x16 tcgen05.mma (the shapes rtkv_flash_prefill issues) with nothing else running: rtkv_debug_mma_rate."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from rtkv import ops

buf = torch.zeros(8, dtype=torch.int64, device="cuda")
names = {0: "SS K-major (Q K^T)", 1: "SS, B MN-major (P in smem)", 2: "TS, B MN-major (P in TMEM)", 3: "8 x SS + 8 x TS", 4: "SS K-major, N = 256", 5: "SS K-major, four accumulators in turn", 6: "SS K-major, descriptors precomputed, unrolled x8"}
n = 2048
for ctas in (148,):
    for mode in range(7):
        for _ in range(2):
            ops.check(ops.lib.rtkv_debug_mma_rate(mode, n, ctas, buf.data_ptr(), None))
            torch.cuda.synchronize()
        c = buf.cpu().tolist()
        print(json.dumps({"ctas": ctas, "mode": names[mode], "issue_cycles_per_mma": round(c[0] / n, 1), "retire_cycles_per_mma": round(c[1] / n, 1)}))
