"""Multi-GPU sharding of the compression path (one process per GPU, torch.distributed).

batch sharding   every reduction of the path is per batch row -> no collective at all; each rank
                 compresses its own sequences (BASELINE configs[4], bench.py --gpus N).
KV-head sharding rank r holds heads [r*H/R, (r+1)*H/R) of Q/K/V and of the attention
                 probabilities (BASELINE configs[2]).  Two small exchanges per layer:
                   1. all-gather of the scalar per-head-group partial sums [B,S,8/R] -> [B,S,8];
                      every rank adds the groups in the same fixed order, so scores, labels and the
                      selection are bit-identical on all ranks and to the single-GPU result;
                   2. MIN/MAX all-reduce of the kept rows' local (kmin,kmax,vmin,vmax) [B,S',4]:
                      the reference's per-token scale spans all heads of the row.
                 Both messages are latency-bound (<= 1 MiB at S = 32768); NCCL over NVLink/NVSwitch.
The collectives are the only torch.distributed calls; with backend "gloo" and CPU tensors the same
functions run on the host, which is how tests/test_dist_cpu.py covers the exchange logic.
"""
import torch
import torch.distributed as dist

N_GROUPS = 8


def batch_shard(batch_size, rank, world):
    """Contiguous batch rows of `rank` (no collective needed for this mode)."""
    per = (batch_size + world - 1) // world
    lo = min(batch_size, rank * per)
    return slice(lo, min(batch_size, lo + per))


def head_shard(num_heads, rank, world):
    """(first head, last head + 1, group_size) of `rank`; the 8 head groups of rtkv_prompt_mass
    must split evenly over the ranks."""
    if num_heads % N_GROUPS or N_GROUPS % world:
        raise ValueError(f"head sharding needs num_heads % 8 == 0 and world_size in (1,2,4,8); got H={num_heads}, R={world}")
    per = num_heads // world
    return rank * per, (rank + 1) * per, num_heads // N_GROUPS


def gather_partials(partials_local, group=None):
    """[B,S,8/R] per rank -> [R,B,S,8/R] on every rank (one all_gather_into_tensor; rank-major is
    head order, ops.mass_from_partials(..., rank_major=True) consumes it without a transpose)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partials_local.unsqueeze(0)
    shape = tuple(partials_local.shape)
    out = torch.empty((world * shape[0],) + shape[1:], dtype=partials_local.dtype, device=partials_local.device)
    dist.all_gather_into_tensor(out, partials_local.contiguous(), group=group)   # concatenation along dim 0
    return out.view((world,) + shape)


def reduce_minmax(minmax_negmax, group=None):
    """[B,S',4] = (min0, -max0, min1, -max1) per rank -> element-wise MIN over the ranks, in place.
    (ops.row_minmax(..., negate_max=True) produces this layout; negation is exact, so MIN of the
    negated maxima is the negated MAX: one collective, bit-exact.)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(minmax_negmax, op=dist.ReduceOp.MIN, group=group)
    return minmax_negmax


class HeadShardedCompressor:
    """compress_layer_kv_cache for a rank that holds a contiguous slice of the KV heads.

    Returns this rank's slice of the compressed K/V ([B,S',HD/R]) and the (replicated) Selection.
    Concatenating the ranks' slices along the last dim reproduces the single-GPU output bit for bit.
    """

    def __init__(self, compressor, num_heads, group=None):
        self.compressor = compressor
        self.num_heads = num_heads
        self.group = group

    def compress_layer_kv_cache(self, key_local, value_local, attention_local, layer_idx):
        from . import ops
        B, S, _ = key_local.shape
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        gs = self.num_heads // N_GROUPS
        comp = self.compressor
        P = comp.prompt_length(S)
        params, _ = comp.layer_params(layer_idx, S, ops.F32)
        partials = ops.prompt_mass_partials(attention_local, P, gs)
        mass = ops.mass_from_partials(gather_partials(partials, self.group), self.num_heads, rank_major=True)
        sel = ops.score_select(params, B, S, key_local.device, mass=mass)
        layer, _rows = sel.read_stats()                       # S' is a shape: one sync
        max_sel = int(layer["max_selected"])
        mm = reduce_minmax(ops.row_minmax(key_local, value_local, sel, max_sel, negate_max=True), self.group)
        k_sel, v_sel = ops.quant_gather_minmax(key_local, value_local, sel, sel.labels, comp._bits(), max_sel, mm,
                                               negated_max=True)
        return k_sel, v_sel, sel

    def compress_layer_async(self, key_local, value_local, attention_local, layer_idx, out_keys, out_values):
        """Synchronisation-free variant (fixed capacity S rows, like compress_layer_kv_cache_async):
        the two exchanges run on the full [B,S,.] extent, nothing is read back to the host."""
        from . import ops
        B, S, _ = key_local.shape
        gs = self.num_heads // N_GROUPS
        comp = self.compressor
        params, _ = comp.layer_params(layer_idx, S, ops.F32)
        partials = ops.prompt_mass_partials(attention_local, comp.prompt_length(S), gs)
        mass = ops.mass_from_partials(gather_partials(partials, self.group), self.num_heads, rank_major=True)
        sel = ops.score_select(params, B, S, key_local.device, mass=mass)
        mm = reduce_minmax(ops.row_minmax(key_local, value_local, sel, S, negate_max=True), self.group)
        ops.quant_gather_minmax(key_local, value_local, sel, sel.labels, comp._bits(), S, mm, zero_pad=False,
                                out0=out_keys, out1=out_values, negated_max=True)
        return sel
