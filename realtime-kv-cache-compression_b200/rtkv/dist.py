"""Multi-GPU sharding of the compression path (one process per GPU, torch.distributed).

batch sharding   every reduction of the path is per batch row -> no collective at all; each rank
                 compresses its own sequences (BASELINE configs[4], bench.py --workload c5).
KV-head sharding rank r holds heads [r*H/R, (r+1)*H/R) of Q/K/V and of the attention
                 probabilities (BASELINE configs[2]).  Two reductions of the reference span all heads of a
                 token: the head mean of the prompt mass (token_importance.py:45) and the per-token min/max of
                 the whole [H*D] row (dynamic_quantization.py:183-184).  Three schemes, identical bits:
                   "fused" (default)  ONE exchange per layer, inside the kernels: rtkv_shard_stats streams the
                      local A/K/V once and peer-stores every token's (kmin,kmax,vmin,vmax, mass partials) record
                      into every rank's symmetric-memory buffer over NVLink; rtkv_shard_combine waits for the R
                      epoch flags and reduces the records in rank order.  No NCCL call on the data path.
                   "collective"  all_gather_into_tensor of the mass partials, then (after the selection) row
                      min/max of the kept rows + ONE MIN all-reduce, then quantise (two NCCL calls per layer).
                   "p2p"  the collective all-gather + rtkv_quant_gather_p2p (row min/max exchanged through peer
                      mailboxes inside the quantise kernel, K/V read once, latency bound).
                 Every rank adds the partials in the same fixed order, so scores, labels and the selection are
                 bit-identical on all ranks and to the single-GPU result.
The NCCL collectives of the "collective" scheme are the only torch.distributed data-path calls; with backend
"gloo" and CPU tensors the same functions run on the host, which is how tests/test_dist_cpu.py covers them.
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
        self.state = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                       # every mailbox is zeroed before anybody writes

    def fits(self, batch, rows):
        return batch * rows <= self.batch * self.rows

    def error(self):
        """True if a kernel gave up waiting for a peer's words (synchronises)."""
        return bool(int(self.state[2].item()))


class ShardExchange:
    """Symmetric (peer-mapped) exchange buffers + local state words for rtkv_shard_stats / rtkv_shard_combine.
    torch's symmetric memory is the plumbing (allocation + the peers' device addresses over NVLink P2P); the
    exchange itself happens inside the rtkv kernels."""

    def __init__(self, batch, seq_len, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.batch, self.seq_len = batch, seq_len
        nbytes = ops.shard_exchange_bytes(batch, seq_len, self.world)
        self.buf = symm_mem.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.local_ptr = self.ptrs[self.rank]
        self.state = torch.zeros(4, dtype=torch.int32, device=device)
        self.mass = torch.empty(batch, seq_len, dtype=torch.float32, device=device)
        self.minmax = torch.empty(batch, seq_len, 4, dtype=torch.float32, device=device)
        self.concurrent = True                         # one process per rank: all ranks run the same kernel at the same time
        self.mass_range = ops.new_mass_range(batch, device)   # reduced by the exchange kernel, consumed + reset by K2
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                       # every buffer is zeroed before anybody writes

    def fits(self, batch, seq_len):
        return (batch, seq_len) == (self.batch, self.seq_len)

    def error(self):
        """True if a combine gave up waiting for a peer (synchronises)."""
        return bool(int(self.state[2].item()))


class LocalExchange:
    """`world` ranks emulated on ONE device (tests, single-GPU profiling): plain device buffers instead of
    symmetric memory; rank r uses .view(r).  Call the producer for every rank before any consumer."""

    class _View:
        pass

    def __init__(self, batch, seq_len, world, device):
        from . import ops
        nbytes = ops.shard_exchange_bytes(batch, seq_len, world)
        self.world = world
        self.bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device) for _ in range(world)]
        self.states = [torch.zeros(4, dtype=torch.int32, device=device) for _ in range(world)]
        self.ptrs = [int(b.data_ptr()) for b in self.bufs]
        self.batch, self.seq_len, self.device = batch, seq_len, device

    def view(self, rank):
        v = LocalExchange._View()
        v.rank, v.world, v.ptrs, v.local_ptr, v.state = rank, self.world, self.ptrs, self.ptrs[rank], self.states[rank]
        v.mass = torch.empty(self.batch, self.seq_len, dtype=torch.float32, device=self.device)
        v.minmax = torch.empty(self.batch, self.seq_len, 4, dtype=torch.float32, device=self.device)
        v.fits = lambda b, s: (b, s) == (self.batch, self.seq_len)
        v.concurrent = False      # ranks run one after the other here: the one-kernel scheme would wait for itself
        return v


class HeadShardedCompressor:
    """compress_layer_kv_cache for a rank that holds a contiguous slice of the KV heads.

    Returns this rank's slice of the compressed K/V ([B,S',HD/R]) and the (replicated) Selection.
    Concatenating the ranks' slices along the last dim reproduces the single-GPU output bit for bit.
    mode: "fused" (one in-kernel exchange per layer, default), "collective" (two NCCL calls per layer),
    "p2p" (NCCL all-gather + min/max mailboxes inside the quantise kernel).  exchange: a ShardExchange /
    LocalExchange view to use instead of allocating one (tests)."""

    def __init__(self, compressor, num_heads, group=None, p2p=False, mode=None, exchange=None):
        self.compressor = compressor
        self.num_heads = num_heads
        self.group = group
        self.mode = mode if mode is not None else ("p2p" if p2p else "fused")
        if self.mode not in ("fused", "collective", "p2p"):
            raise ValueError(f"unknown head-sharding mode {self.mode!r}")
        self.p2p = self.mode == "p2p"
        self.mailbox = None
        self.exchange = exchange
        self._range_pending = False

    def _mailbox(self, batch, rows, device):
        if self.mailbox is None or not self.mailbox.fits(batch, rows):
            self.mailbox = PeerMailbox(batch, rows, device, self.group)
        return self.mailbox

    def _exchange(self, batch, seq_len, device):
        if self.exchange is None or not self.exchange.fits(batch, seq_len):
            self.exchange = ShardExchange(batch, seq_len, device, self.group)
        return self.exchange

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

    # ---- fused scheme: stats + push (+ combine) -> select -> quantise (3 or 4 launches, no NCCL) ----
    def fused_mass_minmax(self, key_local, value_local, attention_local):
        """mass [B,S] and whole-row min/max [B,S,4] of this layer.  Real multi-GPU: ONE kernel (statistics, peer
        stores, per-CTA flags, reduction: rtkv_shard_stats_combine); shapes that kernel cannot take, and ranks emulated
        in one process, use the two-kernel scheme (rtkv_shard_stats + rtkv_shard_combine)."""
        from . import ops
        from ._lib import RtkvError
        B, S, _ = key_local.shape
        x = self._exchange(B, S, key_local.device)
        P, gs = self.compressor.prompt_length(S), self.num_heads // N_GROUPS
        if getattr(x, "concurrent", False) and not getattr(self, "_two_kernels", False):
            try:
                if self._range_pending:          # the previous layer's selection never consumed (and reset) the range
                    x.mass_range.copy_(ops.new_mass_range(B, key_local.device))
                mm = ops.shard_stats_combine(attention_local, key_local, value_local, P, gs, x.ptrs, x.rank, x.world,
                                             x.state, self.num_heads, x.mass, x.minmax, mass_range=x.mass_range)
                self._range_pending = True
                return mm
            except RtkvError as e:
                if e.code != -2:                     # RTKV_EUNSUPPORTED: same verdict on every rank (same shapes)
                    raise
                self._two_kernels = True
        ops.shard_stats(attention_local, key_local, value_local, P, gs, x.ptrs, x.rank, x.world, x.state)
        return ops.shard_combine(x.local_ptr, x.state, B, S, self.num_heads, x.world, x.mass, x.minmax)

    def _fused_select(self, key_local, mass, layer_idx):
        from . import ops
        B, S, _ = key_local.shape
        params, _ = self.compressor.layer_params(layer_idx, S, ops.F32)
        if self._range_pending:                  # the one-kernel exchange also reduced the row's mass range
            self._range_pending = False
            return ops.score_select(params, B, S, key_local.device, mass=mass, mass_range=self.exchange.mass_range)
        return ops.score_select(params, B, S, key_local.device, mass=mass)

    def fused_publish(self, key_local, value_local, attention_local):
        """Producer half (rtkv_shard_stats).  Separate from fused_finish so that a single-process emulation can
        run every rank's producer before any consumer."""
        from . import ops
        B, S, _ = key_local.shape
        x = self._exchange(B, S, key_local.device)
        ops.shard_stats(attention_local, key_local, value_local, self.compressor.prompt_length(S),
                        self.num_heads // N_GROUPS, x.ptrs, x.rank, x.world, x.state)
        return x

    def fused_select(self, key_local, layer_idx):
        from . import ops
        B, S, _ = key_local.shape
        x = self.exchange
        mass, mm = ops.shard_combine(x.local_ptr, x.state, B, S, self.num_heads, x.world, x.mass, x.minmax)
        params, _ = self.compressor.layer_params(layer_idx, S, ops.F32)
        return ops.score_select(params, B, S, key_local.device, mass=mass), mm

    def _fused_quantise(self, key_local, value_local, sel, mm, rows, zero_pad, out_keys=None, out_values=None):
        from . import ops
        return ops.quant_gather_token_minmax(key_local, value_local, sel, sel.labels, self.compressor._bits(), rows, mm,
                                             zero_pad=zero_pad, out0=out_keys, out1=out_values)

    def compress_layer_kv_cache(self, key_local, value_local, attention_local, layer_idx):
        if self.mode == "fused":
            mass, mm = self.fused_mass_minmax(key_local, value_local, attention_local)
            sel = self._fused_select(key_local, mass, layer_idx)
            layer, _rows = sel.read_stats()
            k_sel, v_sel = self._fused_quantise(key_local, value_local, sel, mm, int(layer["max_selected"]), True)
            return k_sel, v_sel, sel
        sel = self._select(key_local, attention_local, layer_idx)
        layer, _rows = sel.read_stats()                       # S' is a shape: one sync
        max_sel = int(layer["max_selected"])
        k_sel, v_sel = self._quantise(key_local, value_local, sel, max_sel, True)
        return k_sel, v_sel, sel

    def compress_layer_async(self, key_local, value_local, attention_local, layer_idx, out_keys, out_values):
        """Synchronisation-free variant (fixed capacity S rows, like compress_layer_kv_cache_async):
        the exchanges run on the full [B,S,.] extent, nothing is read back to the host."""
        if self.mode == "fused":
            mass, mm = self.fused_mass_minmax(key_local, value_local, attention_local)
            sel = self._fused_select(key_local, mass, layer_idx)
            self._fused_quantise(key_local, value_local, sel, mm, key_local.shape[1], False, out_keys, out_values)
            return sel
        sel = self._select(key_local, attention_local, layer_idx)
        self._quantise(key_local, value_local, sel, key_local.shape[1], False, out_keys, out_values)
        return sel
