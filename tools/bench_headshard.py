#!/usr/bin/env python
"""Stage timings of the head-sharded layer (torchrun, NCCL): where does the time go?"""
import os, sys, json, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "realtime-kv-cache-compression_b200"))
import torch, torch.distributed as dist
import rtkv
from rtkv import ops, dist as rd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, S, D, P = 1, 32, 32768, 128, 128
hl, w = H // world, H * D // world
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand(B, hl, S, P, generator=g, device=dev) / P
K = torch.randn(B, S, w, generator=g, device=dev)
V = torch.randn(B, S, w, generator=g, device=dev)
ok, ov = torch.empty_like(K), torch.empty_like(V)
cfg = rtkv.CompressionConfig(high_precision_bits=8, medium_precision_bits=4, low_precision_bits=2)
comp = rtkv.RealTimePrefillCompressor(cfg)
params, _ = comp.layer_params(5, S, ops.F32)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


res = {}
parts = ops.prompt_mass_partials(A, P, H // 8)
res["k1_partials_us"] = timeit(lambda: ops.prompt_mass_partials(A, P, H // 8))
res["all_gather_list_us"] = timeit(lambda: rd.gather_partials(parts))
flat = torch.empty(world, *parts.shape, device=dev)
res["all_gather_into_tensor_us"] = timeit(lambda: dist.all_gather_into_tensor(flat, parts))
full = rd.gather_partials(parts)
mass = ops.mass_from_partials(full, H, rank_major=True)
res["mass_from_partials_us"] = timeit(lambda: ops.mass_from_partials(full, H, rank_major=True))
sel = ops.score_select(params, B, S, dev, mass=mass)
res["k2_us"] = timeit(lambda: ops.score_select(params, B, S, dev, mass=mass))
mm = ops.row_minmax(K, V, sel, S, negate_max=True)
res["row_minmax_us"] = timeit(lambda: ops.row_minmax(K, V, sel, S, negate_max=True))
res["all_reduce_min_us"] = timeit(lambda: dist.all_reduce(mm, op=dist.ReduceOp.MIN))
res["k3_minmax_us"] = timeit(lambda: ops.quant_gather_minmax(K, V, sel, sel.labels, (2, 4, 8), S, mm, zero_pad=False, out0=ok, out1=ov, negated_max=True))
res["k3_plain_us"] = timeit(lambda: ops.quant_gather(K, V, sel, sel.labels, (2, 4, 8), S, zero_pad=False, out0=ok, out1=ov))
mb = rd.PeerMailbox(B, S, dev)
res["k3_p2p_fused_us"] = timeit(lambda: ops.quant_gather_p2p(K, V, sel, sel.labels, (2, 4, 8), S, mb.ptrs, mb.rank, mb.world,
                                                             mb.state, zero_pad=False, out0=ok, out1=ov))
if rank == 0:
    print(json.dumps({k: round(v, 1) for k, v in res.items()}), flush=True)
threading.Timer(15.0, lambda: os._exit(0)).start()
dist.destroy_process_group()
os._exit(0)
