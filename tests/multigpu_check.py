#!/usr/bin/env python
"""Real multi-GPU check of KV-head sharding (run under torchrun on R = 2, 4 or 8 GPUs, NCCL):
every rank compresses its head slice with each exchange scheme of rtkv/dist.py ("fused": one in-kernel
exchange over symmetric memory, "collective": two NCCL calls, "p2p": min/max mailboxes); the concatenated
result must equal the single-GPU compress_layer_kv_cache bit for bit.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import rtkv  # noqa: E402
from rtkv import dist as rd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok_all = True
    for (B, H, S, D, layer) in [(2, 16, 700, 32, 3), (1, 32, 4096, 128, 25), (1, 32, 9000, 64, 12)]:
        g = torch.Generator(device="cuda").manual_seed(1234)          # identical data on every rank
        HD = H * D
        P = max(1, min(S // 5, 128))
        A = torch.rand(B, H, S, P, generator=g, device="cuda")
        A = A / A.sum(-1, keepdim=True) * torch.rand(B, H, S, 1, generator=g, device="cuda")
        K = torch.randn(B, S, HD, generator=g, device="cuda")
        V = torch.randn(B, S, HD, generator=g, device="cuda") * 2 + 0.5
        cfg = rtkv.CompressionConfig(num_attention_heads=H, hidden_size=HD, high_precision_bits=8,
                                     medium_precision_bits=4, low_precision_bits=2)
        comp = rtkv.RealTimePrefillCompressor(cfg)
        ids = torch.zeros(B, S, dtype=torch.long, device="cuda")
        k_ref, v_ref, info = comp.compress_layer_kv_cache(K, V, A, ids, layer)
        h0, h1, gs = rd.head_shard(H, rank, world)
        w = HD // world
        for mode in ("fused", "collective", "p2p"):
            sharded = rd.HeadShardedCompressor(comp, H, mode=mode)
            Kl = K[..., rank * w:(rank + 1) * w].contiguous()
            Vl = V[..., rank * w:(rank + 1) * w].contiguous()
            Al = A[:, h0:h1].contiguous()
            ok = True
            for rep in range(5):                                   # repeated calls exercise epoch / parity reuse
                k_loc, v_loc, sel = sharded.compress_layer_kv_cache(Kl, Vl, Al, layer)
                ks = [torch.empty_like(k_loc) for _ in range(world)]
                vs = [torch.empty_like(v_loc) for _ in range(world)]
                dist.all_gather(ks, k_loc)
                dist.all_gather(vs, v_loc)
                ok = ok and (torch.equal(torch.cat(ks, -1), k_ref) and torch.equal(torch.cat(vs, -1), v_ref)
                             and torch.equal(sel.mask.view(torch.bool), info["propagation_info"]["selection_mask"]))
            okb, ovb = torch.zeros(B, S, w, device="cuda"), torch.zeros(B, S, w, device="cuda")
            sharded.compress_layer_async(Kl, Vl, Al, layer, okb, ovb)
            ms = k_ref.shape[1]
            ok2 = (torch.equal(okb[:, :ms], k_loc) and torch.equal(ovb[:, :ms], v_loc)) if B == 1 else True
            if mode == "fused":
                ok2 = ok2 and not sharded.exchange.error()
            t = torch.tensor([int(ok and ok2)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"head-sharded R={world} B={B} H={H} S={S} mode={mode}: kept {k_ref.shape[1]} ->",
                      "PASS" if int(t) else "FAIL", flush=True)
            ok_all = ok_all and bool(int(t))
    import threading
    threading.Timer(15.0, lambda: os._exit(0 if ok_all else 1)).start()
    dist.destroy_process_group()
    os._exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
