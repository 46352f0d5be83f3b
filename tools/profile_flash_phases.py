"""Where do the cycles of flash_prefill_kernel's tile loop go?  Runs the instrumented build (rtkv_debug_flash_profile) and prints,
for CTA 0 (the last = heaviest query tile of head 0), the average cycles per key tile in each phase of one softmax warp of each
warpgroup, and how long the MMA thread waited for P.  Not a benchmark: the clock reads perturb the loop a little."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from rtkv import ops


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # 0 (default kernel) or 3 (alternating key tiles)
    ops.lib.rtkv_debug_flash_variant(variant)
    H = 32
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(1, H, S, 128, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    if len(sys.argv) > 3:
        buf[63] = 1          # variant 4: the K producer also waits for each of its loads (latency figure; perturbs the run)
    ops.lib.rtkv_debug_flash_profile(buf.data_ptr())
    try:
        for _ in range(3):
            ops.flash_prefill(q, k, v, causal=True, prompt_len=128)
        torch.cuda.synchronize()
    finally:
        ops.lib.rtkv_debug_flash_profile(None)
        ops.lib.rtkv_debug_flash_variant(-1)
    c = buf.cpu().tolist()
    out = {"S": S, "variant": variant}
    for g_ in range(2):
        T = c[g_ * 16 + 8]
        # counters: [0] = time between the end of one tile and the top of the next (incl. loop overhead), [1] wait_S, [2] ld, [3] max,
        # [4] barrier + rescale check, [5] wait for the P buffer, [6] exp/pack/store, [7] fence + arrive
        names = ["loop_top", "wait_S", "tmem_ld+release_S", "row_max", "exchange_barrier+rescale_check", "wait_P_buffer", "exp+pack+store", "fence+arrive"]
        out[f"warpgroup{g_}"] = {"tiles": T, **{n: round(c[g_ * 16 + i] / max(T, 1), 1) for i, n in enumerate(names)},
                                 "sum_per_tile": round(sum(c[g_ * 16:g_ * 16 + 8]) / max(T, 1), 1)}
    T = c[35]
    out["mma_thread"] = {"tiles": T, "wait_P_per_tile": round(c[32] / max(T, 1), 1), "issue_next_QK_per_tile": round(c[34] / max(T, 1), 1),
                         "loop_per_tile": round(c[33] / max(T, 1), 1)}
    if variant == 4:          # two issuing threads: [34] = the QK thread's loop, [33] = the PV thread's
        out["mma_thread"] = {"tiles": T, "QK_thread_loop_per_tile": round(c[34] / max(T, 1), 1), "PV_thread_loop_per_tile": round(c[33] / max(T, 1), 1)}
    if variant == 4:
        names = ["wait_K", "wait_S_buffer", "issue_QK", "wait_P", "wait_V", "issue_PV"]
        out["mma_thread"].update({n: round(c[36 + i] / max(T, 1), 1) for i, n in enumerate(names)})
        out["k_producer"] = {"wait_free_slot_per_tile": round(c[44] / max(T, 1), 1), "load_latency_per_tile": round(c[45] / max(T, 1), 1)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
