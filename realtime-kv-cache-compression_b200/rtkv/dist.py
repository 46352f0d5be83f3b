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


class PeerMailbox:
    """Symmetric (peer-mapped) mailbox for the fused min/max exchange of rtkv_quant_gather_p2p.
    torch's symmetric memory is the plumbing: it allocates the buffer and hands out the peers'
    device addresses (NVLink P2P); the exchange itself happens inside the rtkv kernel."""

    def __init__(self, batch, rows, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.batch, self.rows = batch, rows
        nbytes = int(_lib.lib.rtkv_p2p_mailbox_bytes(int(batch), int(rows), int(self.world)))
        if nbytes <= 0:
            raise ValueError("bad mailbox shape")
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                       # every mailbox is zeroed before anybody writes

    def fits(self, batch, rows):
        return batch * rows <= self.batch * self.rows


class HeadShardedCompressor:
    """compress_layer_kv_cache for a rank that holds a contiguous slice of the KV heads.

    Returns this rank's slice of the compressed K/V ([B,S',HD/R]) and the (replicated) Selection.
    Concatenating the ranks' slices along the last dim reproduces the single-GPU output bit for bit.
    p2p=True uses the fused kernel (row min/max exchanged over NVLink inside the quantise kernel,
    K/V read once); p2p=False the two-pass scheme with a MIN all-reduce.
    """

    def __init__(self, compressor, num_heads, group=None, p2p=False):
        self.compressor = compressor
        self.num_heads = num_heads
        self.group = group
        self.p2p = p2p
        self.mailbox = None

    def _mailbox(self, batch, rows, device):
        if self.mailbox is None or not self.mailbox.fits(batch, rows):
            self.mailbox = PeerMailbox(batch, rows, device, self.group)
        return self.mailbox

    def _quantise(self, key_local, value_local, sel, rows, zero_pad, out_keys=None, out_values=None):
        from . import ops
        comp = self.compressor
        if self.p2p:
            mb = self._mailbox(key_local.shape[0], key_local.shape[1], key_local.device)
            return ops.quant_gather_p2p(key_local, value_local, sel, sel.labels, comp._bits(), rows, mb.ptrs, mb.rank,
                                        mb.world, mb.state, zero_pad=zero_pad, out0=out_keys, out1=out_values)
        mm = reduce_minmax(ops.row_minmax(key_local, value_local, sel, rows, negate_max=True), self.group)
        return ops.quant_gather_minmax(key_local, value_local, sel, sel.labels, comp._bits(), rows, mm,
                                       zero_pad=zero_pad, out0=out_keys, out1=out_values, negated_max=True)

    def _select(self, key_local, attention_local, layer_idx):
        from . import ops
        B, S, _ = key_local.shape
        gs = self.num_heads // N_GROUPS
        comp = self.compressor
        params, _ = comp.layer_params(layer_idx, S, ops.F32)
        partials = ops.prompt_mass_partials(attention_local, comp.prompt_length(S), gs)
        mass = ops.mass_from_partials(gather_partials(partials, self.group), self.num_heads, rank_major=True)
        return ops.score_select(params, B, S, key_local.device, mass=mass)

    def compress_layer_kv_cache(self, key_local, value_local, attention_local, layer_idx):
        sel = self._select(key_local, attention_local, layer_idx)
        layer, _rows = sel.read_stats()                       # S' is a shape: one sync
        max_sel = int(layer["max_selected"])
        k_sel, v_sel = self._quantise(key_local, value_local, sel, max_sel, True)
        return k_sel, v_sel, sel

    def compress_layer_async(self, key_local, value_local, attention_local, layer_idx, out_keys, out_values):
        """Synchronisation-free variant (fixed capacity S rows, like compress_layer_kv_cache_async):
        the exchanges run on the full [B,S,.] extent, nothing is read back to the host."""
        sel = self._select(key_local, attention_local, layer_idx)
        self._quantise(key_local, value_local, sel, key_local.shape[1], False, out_keys, out_values)
        return sel
