// head_shard.cu — KV-head sharding with ONE exchange per layer.
//
// A rank holds H/R consecutive heads: its slice of the attention probabilities [B,H/R,S,>=P] and of the
// K/V rows [B,S,(H/R)*D].  Two reductions of the reference span ALL heads of a token:
//   src/compression/token_importance.py:45          A[:, :, :, prompt].mean(1).sum(-1)   (head mean of the prompt mass)
//   src/compression/dynamic_quantization.py:183-184 per-token min/max over the whole [H*D] row
// so every token needs one record from every rank before anything else can happen:
//   record = (kmin, kmax, vmin, vmax, partial[0..8/R))       partial[g] = prompt mass of local head group g
//
// shard_stats_kernel    streams the local A slice and the local K/V rows ONCE, builds the record of every token
//                       (one warp per token, the head-group reduction order of prompt_mass.cuh) and stores it into
//                       the exchange buffer of EVERY rank with peer stores over NVLink (the buffers are symmetric
//                       memory mapped into this process).  The last CTA publishes an epoch flag to every rank
//                       (release at system scope), so the "all-gather" costs no extra launch and no NCCL call.
// shard_combine_kernel  waits for the R epoch flags (acquire), then reduces the R records of every token in rank
//                       order: mass = (p0 + ... + p7)/H exactly like rtkv_mass_from_partials (bit-identical to the
//                       unsharded rtkv_prompt_mass), row min/max over ranks -> [B,S,4] indexed by TOKEN, which
//                       rtkv_quant_gather_token_minmax consumes after the (replicated) selection.
//
// Buffers are double-buffered by epoch parity.  A rank can start epoch e+2 only after its own combine of e+1 has
// seen every rank's flag e+1, i.e. after every rank has finished reading the records of epoch e: two parities are
// enough and no other synchronisation between ranks exists.  The epoch is a device-side counter, so the kernels are
// CUDA-graph replayable.
#include <stdlib.h>

#include "prompt_mass.cuh"

namespace rtkv {

constexpr int kMaxShardCtas = 256;   // per-CTA flags of the fused statistics+combine kernel (grid <= SM count)
constexpr int kXHeader = 256 + 2 * 8 * kMaxShardCtas * 4;   // bytes: flags[2 parities][8 sources] u32 (padded to 256) + CTA flags[2][8][256] u32
constexpr int kShardWarps = 8;

struct ShardParams {
    const void *A;
    int64_t sb, sh, si;
    const void *K, *V;
    int B, S, P, gs, ng, n;           // ng local head groups of gs heads; n = local row width in elements
    int rank, world, rec_vecs, dbg, kv_keep;   // kv_keep: leading token rows whose K/V slices are asked to stay in L2
    //    // dbg (rtkv_debug_shard_variant): 2 no consumer math, 3 no record stores
    unsigned long long xbuf[8];       // exchange buffer of rank s as mapped into this process
    unsigned *state;                  // local: [0] epoch of the last finished call, [1] CTA ticket, [2] error flag
    // fused combine (rtkv_shard_stats_combine): this rank's outputs, NULL = publish only
    float *mass, *minmax;
    float fH;
    unsigned *mass_range;             // optional [B][4] (min key, max key, NaN flag, pad): see rtkv_score_select_ranged
};

__device__ __forceinline__ void st_sys_v4(void *p, float a, float b, float c, float d)
{
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// release/acquire fence at system scope.  __threadfence_system() compiles to the sequentially consistent MEMBAR.SC.SYS;
// the record -> flag hand-off only needs acquire/release ordering (RTKV_SHARD_SC_FENCE=1 at build time restores SC).
__device__ __forceinline__ void fence_sys()
{
#ifdef RTKV_SHARD_SC_FENCE
    __threadfence_system();
#else
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
}
__device__ __forceinline__ void st_sys_u32(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_sys_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg_v4(const void *p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ size_t record_offset(unsigned parity, int world, int src, int64_t rows, int64_t row, int rec_vecs)
{
    return (size_t)kXHeader + ((((size_t)parity * world + src) * (size_t)rows + (size_t)row) * rec_vecs) * 16;
}

// mass and whole-row min/max of one token from the R records of epoch `parity` in this rank's exchange buffer
// == score_key_desc of score_select.cuh (torch.argsort(descending) order: NaN first, -0 == +0): the mass range handed
// to the selection kernel is expressed in its key domain
__device__ __forceinline__ unsigned mass_key_desc(float s)
{
    if (s != s) return 0u;
    unsigned u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;
    const unsigned asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ float combine_token(const char *xlocal, unsigned parity, int world, int gpr, int rec_vecs, int64_t rows,
                                              int64_t t, float fH, float *__restrict__ mass, float *__restrict__ minmax)
{
    float tot = 0.0f;
    float kmn = INFINITY, kmx = -INFINITY, vmn = INFINITY, vmx = -INFINITY;
    bool knan = false, vnan = false;
    for (int r = 0; r < world; ++r) {
        const char *rec = xlocal + record_offset(parity, world, r, rows, t, rec_vecs);
        const float4 m = ld_cg_v4(rec);
        const float4 a = ld_cg_v4(rec + 16);
        float pg[8] = {a.x, a.y, a.z, a.w, 0.0f, 0.0f, 0.0f, 0.0f};
        if (rec_vecs > 2) {
            const float4 c = ld_cg_v4(rec + 32);
            pg[4] = c.x; pg[5] = c.y; pg[6] = c.z; pg[7] = c.w;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
            if (g < gpr) tot = (r == 0 && g == 0) ? pg[g] : __fadd_rn(tot, pg[g]);
        knan = knan || (m.x != m.x); vnan = vnan || (m.z != m.z);
        kmn = fminf(kmn, m.x); kmx = fmaxf(kmx, m.y); vmn = fminf(vmn, m.z); vmx = fmaxf(vmx, m.w);
    }
    const float qnan = __int_as_float(0x7fc00000);
    const float m = __fdiv_rn(tot, fH);
    mass[t] = m;
    float4 o = make_float4(knan ? qnan : kmn, knan ? qnan : kmx, vnan ? qnan : vmn, vnan ? qnan : vmx);
    *reinterpret_cast<float4 *>(minmax + 4 * t) = o;
    return m;
}

// End of a statistics kernel: every record of this rank must be visible at every peer before the epoch flag is.
// Every thread fences its own record stores at system scope (in parallel: one fence latency), the CTA barrier orders
// them before thread 0's ticket, and the LAST CTA fences once more (acquire of the other CTAs' tickets + release of
// everything they ordered before them) and then writes the flags with plain system-scope stores.  One fence per
// thread plus one in the last CTA: a chain of st.release.sys (a system fence each, ~2.3 us apiece) measured 25 us.
__device__ __forceinline__ void publish_epoch(const ShardParams &p, unsigned parity, unsigned epoch, unsigned n_ctas)
{
    fence_sys();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&p.state[1], 1u) == n_ctas - 1) {
            p.state[1] = 0u;
            fence_sys();
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < p.world) st_sys_u32((unsigned *)p.xbuf[s] + parity * 8 + p.rank, epoch);
            ((volatile unsigned *)p.state)[0] = epoch;
        }
    }
}

// Generic path (any P, any strides): one warp per token, loads straight into registers.
// DT = dtype of K/V; the probabilities are float32 (half probabilities round the per-column head mean, which does
// not decompose over ranks: rtkv_prompt_mass_partials has the same restriction).
template <int DT, bool VEC>
__global__ void __launch_bounds__(kShardWarps * 32) shard_stats_generic_kernel(ShardParams p)
{
    constexpr int EPV = Vec16<DT>::kElems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();
    pdl_trigger();
    const unsigned epoch = ((volatile unsigned *)p.state)[0] + 1u;
    const unsigned parity = epoch & 1u;
    const int64_t rows = (int64_t)p.B * p.S;
    const int nvec = p.n / EPV;
    unsigned long long my_dst = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s)
        if (lane == s) my_dst = p.xbuf[s];

    for (int64_t row = (int64_t)blockIdx.x * kShardWarps + warp; row < rows; row += (int64_t)gridDim.x * kShardWarps) {
        // ---- local min/max of the K row slice and of the V row slice ----
        float mm[4];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint4 *src = (const uint4 *)((const char *)(t ? p.V : p.K) + (size_t)row * p.n * Elem<DT>::kBytes);
            float mn = INFINITY, mx = -INFINITY;
            int nan = 0;
            for (int v0 = lane; v0 < nvec; v0 += 4 * 32) {
                uint4 raw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (v0 + 32 * k < nvec) raw[k] = ldg_stream(src + v0 + 32 * k);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (v0 + 32 * k < nvec) {
                        float f[EPV];
                        Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
                    }
                }
            }
            mn = warp_min(mn); mx = warp_max(mx);
            if (__any_sync(0xffffffffu, nan)) mn = mx = __int_as_float(0x7fc00000);   // torch.min/max propagate NaN
            mm[2 * t] = mn; mm[2 * t + 1] = mx;
        }
        // ---- prompt mass of the local head groups (same order as prompt_mass_kernel's partials) ----
        const int b = (int)(row / p.S);
        const int i = (int)(row - (int64_t)b * p.S);
        const float *rowp = (const float *)p.A + b * p.sb + (int64_t)i * p.si;
        float pg[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            pg[g] = 0.0f;
            if (g < p.ng) {
                float lane_acc = 0.0f;
                for (int c0 = 0; c0 < p.P; c0 += 128) {
                    float s[4];
                    group_column_sums<RTKV_F32, VEC>(rowp, p.sh, g * p.gs, (g + 1) * p.gs, c0 + 4 * lane, p.P, nullptr, s);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c0 + 4 * lane + k < p.P) lane_acc = __fadd_rn(lane_acc, s[k]);
                }
                pg[g] = butterfly(lane_acc);
            }
        }
        // ---- one record per destination rank (lane s -> rank s), 16-byte peer stores ----
        if (lane < p.world) {
            char *dst = (char *)my_dst + record_offset(parity, p.world, p.rank, rows, row, p.rec_vecs);
            st_sys_v4(dst, mm[0], mm[1], mm[2], mm[3]);
            st_sys_v4(dst + 16, pg[0], pg[1], pg[2], pg[3]);
            if (p.rec_vecs > 2) st_sys_v4(dst + 32, pg[4], pg[5], pg[6], pg[7]);
        }
    }
    // ---- publish: every record of this rank is out before the flag ----
    publish_epoch(p, parity, epoch, gridDim.x);
}

// ---- staged path: bulk async copies (TMA engine, cp.async.bulk) into a ring of shared-memory stages ----
// The register kernel above is latency bound (one row per warp, three dependent load phases: 1.1 - 3.4 TB/s measured).
// Here a producer thread asks the copy engine for whole BLOCKS of rows (completion counted in bytes on an mbarrier),
// so up to ~200 KB per SM are in flight with no registers held, while 8 consumer warps reduce the block that has landed.
// Same arithmetic and the same reduction order as the generic kernel (and as prompt_mass_kernel): bit-identical records.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// L2 residency: the K/V rows are read again by the quantise kernel a few tens of microseconds later (after the
// exchange and the selection), the probability slice never again.  With 8 ranks a rank's K/V slice of a 32k-token layer
// is 64 MiB - it fits the 126 MB L2 if the 64 MiB of probabilities streaming through do not evict it.
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

struct StagedCfg {
    int rb, stages, stage_bytes, kv_bytes, a_bytes;   // rb rows per block; a_bytes = P*4 bytes of one head's prompt columns
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_sys_v2(void *p, float a, float b)
{
    asm volatile("st.volatile.global.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}

constexpr int kStagedConsumers = 16;                          // consumer warps (one row of the block each)
constexpr int kStagedThreads = (kStagedConsumers + 1) * 32;   // + 1 producer warp
constexpr int kStagedHeader = 1024 + 8 * kStagedConsumers * 48;   // barriers + record staging (8 stages x 16 rows x 48 B)

// order-preserving unsigned image of a float: warp min/max become single REDUX instructions
__device__ __forceinline__ unsigned f2ord_u(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_u(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// One CTA per SM, persistent over blocks of `rb` consecutive token rows of one batch row.  Per block the producer
// thread issues 2 + H_local bulk copies (K rows, V rows: contiguous; per head the rb prompt-column runs of A, which
// are contiguous when A is a [B,H,S,P] slice) - a few LARGE copies, because the copy engine costs ~45 ns per copy
// whatever its size (measured: per-row 512-byte copies ran at 2.5 TB/s).  16 consumer warps share the block's
// 3*rb (row, task) items: K min/max, V min/max (one REDUX each) or the prompt-mass partials of a row.
template <int DT>
__device__ __forceinline__ void shard_stats_staged_body(const ShardParams &p, const StagedCfg &c, const int cta, const int n_ctas)
{
    constexpr int EPV = Vec16<DT>::kElems;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t full0 = smem_u32(smem), empty0 = full0 + 64;               // [stages] each (stages <= 8)
    float *rec_s = reinterpret_cast<float *>(smem + 1024);                     // [stages][rb][4*rec_vecs] record staging
    unsigned char *data = smem + kStagedHeader;
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kStagedConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    const unsigned epoch = ((volatile unsigned *)p.state)[0] + 1u;
    const unsigned parity = epoch & 1u;
    const int64_t rows = (int64_t)p.B * p.S;
    const int nvec = p.n / EPV;
    const int Hl = p.ng * p.gs;
    const int bpb = (p.S + c.rb - 1) / c.rb;                   // blocks per batch row
    const int64_t n_blocks = (int64_t)p.B * bpb;

    if (warp == kStagedConsumers) {
        // ---- producer ----
        if (lane == 0) {
            // kv_keep = number of leading token rows (per batch row) whose K/V slices are asked to STAY in L2: the quantise kernel
            // re-reads the kept rows in ascending order a few tens of microseconds later, so the first rows are the ones worth
            // keeping; everything else (and all of A, which is never read again) streams through with evict-first
            const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
            const uint64_t pol_a = pol_stream;
            int64_t k = 0;
            for (int64_t blk = cta; blk < n_blocks; blk += n_ctas, ++k) {
                const int st = (int)(k % c.stages);
                mbar_wait(empty0 + 8 * st, (uint32_t)(((k / c.stages) & 1) ^ 1));
                const int b = (int)(blk / bpb);
                const int i0 = (int)(blk - (int64_t)b * bpb) * c.rb;
                const int nr = min(c.rb, p.S - i0);
                const uint32_t bar = full0 + 8 * st;
                const uint32_t dst = smem_u32(data + (size_t)st * c.stage_bytes);
                mbar_expect_tx(bar, (uint32_t)nr * (2u * c.kv_bytes + (uint32_t)Hl * c.a_bytes));
                const size_t row0 = (size_t)b * p.S + i0;
                const uint64_t pol_kv = i0 < p.kv_keep ? pol_keep : pol_stream;
                bulk_g2s(dst, (const char *)p.K + row0 * c.kv_bytes, nr * c.kv_bytes, bar, pol_kv);
                bulk_g2s(dst + c.rb * c.kv_bytes, (const char *)p.V + row0 * c.kv_bytes, nr * c.kv_bytes, bar, pol_kv);
                const float *a0 = (const float *)p.A + b * p.sb + (int64_t)i0 * p.si;
                for (int h = 0; h < Hl; ++h)
                    bulk_g2s(dst + 2 * c.rb * c.kv_bytes + h * c.rb * c.a_bytes, a0 + (int64_t)h * p.sh, nr * c.a_bytes, bar, pol_a);
            }
        }
    } else {
        // ---- consumers ----
        const int a_vecs = c.a_bytes / 16;                     // float4 per head row (P/4)
        int64_t k = 0;
        for (int64_t blk = cta; blk < n_blocks; blk += n_ctas, ++k) {
            const int st = (int)(k % c.stages);
            const int b = (int)(blk / bpb);
            const int i0 = (int)(blk - (int64_t)b * bpb) * c.rb;
            const int nr = min(c.rb, p.S - i0);
            mbar_wait(full0 + 8 * st, (uint32_t)((k / c.stages) & 1));
            const unsigned char *sd = data + (size_t)st * c.stage_bytes;
            // items = (row of the block, task): task 0 / 1 = min/max of the K / V row slice, 2 = prompt-mass partials;
            // each item writes its part of the token's record into the stage's record staging area
            float *rec_st = rec_s + (size_t)st * kStagedConsumers * 12;
            for (int item = warp; item < 3 * nr && p.dbg != 2; item += kStagedConsumers) {
                const int r = item / 3, task = item - 3 * r;
                float *rs = rec_st + r * 4 * p.rec_vecs;
                if (task < 2) {
                    // per lane fminf/fmaxf, across lanes ONE REDUX per value on the order-preserving unsigned image (same
                    // value as the shuffle tree of the generic kernel; only the sign of a zero extreme could differ, and
                    // -0.0 == +0.0 everywhere downstream: scale, zero-point, mx == mn)
                    const uint4 *src = reinterpret_cast<const uint4 *>(sd + (size_t)task * c.rb * c.kv_bytes + (size_t)r * c.kv_bytes);
                    float mn = INFINITY, mx = -INFINITY;
                    int nan = 0;
                    for (int v0 = lane; v0 < nvec; v0 += 4 * 32) {
                        uint4 raw[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (v0 + 32 * q < nvec) raw[q] = src[v0 + 32 * q];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (v0 + 32 * q < nvec) {
                                float f[EPV];
                                Vec16<DT>::unpack(raw[q], f);
#pragma unroll
                                for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
                            }
                        }
                    }
                    mn = ord2f_u(__reduce_min_sync(0xffffffffu, f2ord_u(mn)));
                    mx = ord2f_u(__reduce_max_sync(0xffffffffu, f2ord_u(mx)));
                    if (__any_sync(0xffffffffu, nan)) mn = mx = __int_as_float(0x7fc00000);   // torch.min/max propagate NaN
                    if (lane == 0) { rs[2 * task] = mn; rs[2 * task + 1] = mx; }
                } else {
                    // heads of a group added in order per column, the lane's 4 columns in order, xor butterfly
                    // (== group_column_sums + butterfly of prompt_mass.cuh)
                    const float4 *ap = reinterpret_cast<const float4 *>(sd + 2 * (size_t)c.rb * c.kv_bytes);
                    float pg[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        pg[g] = 0.0f;
                        if (g < p.ng) {
                            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, lane_acc = 0.0f;
                            if (lane < a_vecs) {
                                for (int h = g * p.gs; h < (g + 1) * p.gs; ++h) {
                                    const float4 v = ap[((size_t)h * c.rb + r) * a_vecs + lane];
                                    s0 = __fadd_rn(s0, v.x); s1 = __fadd_rn(s1, v.y); s2 = __fadd_rn(s2, v.z); s3 = __fadd_rn(s3, v.w);
                                }
                                lane_acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(lane_acc, s0), s1), s2), s3);
                            }
                            pg[g] = butterfly(lane_acc);
                        }
                    }
                    if (lane == 0) {
                        rs[4] = pg[0]; rs[5] = pg[1]; rs[6] = pg[2]; rs[7] = pg[3];
                        if (p.rec_vecs > 2) { rs[8] = pg[4]; rs[9] = pg[5]; rs[10] = pg[6]; rs[11] = pg[7]; }
                    }
                }
            }
            // The block's records are contiguous in every rank's buffer (nr x 32 bytes): once all items are done, warp w
            // sends them to rank w as 16-byte vectors, a warp-wide store covering whole 128-byte lines.  (Per-item 8/16-byte
            // peer stores were 688k tiny NVLink packets per layer and cost ~30 us per layer on 8 GPUs.)
            asm volatile("bar.sync 1, %0;" ::"n"(kStagedConsumers * 32) : "memory");
            if (warp < p.world && p.dbg < 2) {
                unsigned long long base = 0;
#pragma unroll
                for (int sdst = 0; sdst < 8; ++sdst)
                    if (warp == sdst) base = p.xbuf[sdst];
                const int64_t row0 = (int64_t)b * p.S + i0;
                float4 *dst = reinterpret_cast<float4 *>((char *)base + record_offset(parity, p.world, p.rank, rows, row0, p.rec_vecs));
                const float4 *srcv = reinterpret_cast<const float4 *>(rec_st);
                for (int v = lane; v < nr * p.rec_vecs; v += 32) {
                    const float4 x = srcv[v];
                    st_sys_v4(dst + v, x.x, x.y, x.z, x.w);
                }
            }
            __syncwarp();                                  // every lane is done with this stage's shared memory
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
        }
        if (p.mass) {
            // ---- fused combine: CTA c of every rank handles the same token blocks, so this CTA only needs the records
            // of the peers' CTA c.  Point-to-point: fence my record stores, raise my CTA flag at every rank, wait for the
            // R flags of CTA c here, then reduce my own tokens.  No grid-wide flag, no second kernel.
            fence_sys();
            asm volatile("bar.sync 1, %0;" ::"n"(kStagedConsumers * 32) : "memory");
            if (warp == 0 && lane < p.world) {
                unsigned long long base = 0;
#pragma unroll
                for (int sdst = 0; sdst < 8; ++sdst)
                    if (lane == sdst) base = p.xbuf[sdst];
                st_sys_u32((unsigned *)(base + 256) + ((size_t)parity * 8 + p.rank) * kMaxShardCtas + cta, epoch);
                unsigned long long mine = 0;
#pragma unroll
                for (int sdst = 0; sdst < 8; ++sdst)
                    if (p.rank == sdst) mine = p.xbuf[sdst];
                const unsigned *flag = (const unsigned *)(mine + 256) + ((size_t)parity * 8 + lane) * kMaxShardCtas + cta;
                const long long t0 = clock64();
                while (ld_sys_u32(flag) != epoch) {
                    if (clock64() - t0 > 40000000000LL) { atomicExch(&p.state[2], 1u); break; }     // ~20 s: flag it, never trap
                }
                fence_sys();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kStagedConsumers * 32) : "memory");
            unsigned long long mine = 0;
#pragma unroll
            for (int sdst = 0; sdst < 8; ++sdst)
                if (p.rank == sdst) mine = p.xbuf[sdst];
            const int per_blk = c.rb;
            const int64_t my_blocks = cta < n_blocks ? (n_blocks - cta + n_ctas - 1) / n_ctas : 0;
            unsigned rk_min = 0xffffffffu, rk_max = 0u, rk_nan = 0u;       // this thread's share of the row's mass range
            int rk_b = -1;
            for (int64_t idx = threadIdx.x; idx < my_blocks * per_blk; idx += kStagedConsumers * 32) {
                const int64_t blk = cta + (idx / per_blk) * n_ctas;
                const int r = (int)(idx % per_blk);
                const int b = (int)(blk / bpb);
                const int i = (int)(blk - (int64_t)b * bpb) * c.rb + r;
                if (i < p.S) {
                    const float m = combine_token((const char *)mine, parity, p.world, p.ng, p.rec_vecs, rows, (int64_t)b * p.S + i, p.fH,
                                                  p.mass, p.minmax);
                    if (p.mass_range) {
                        if (b != rk_b && rk_b >= 0) {                  // batch row changed: flush (rare)
                            atomicMin(p.mass_range + 4 * rk_b, rk_min); atomicMax(p.mass_range + 4 * rk_b + 1, rk_max);
                            if (rk_nan) atomicOr(p.mass_range + 4 * rk_b + 2, 1u);
                            rk_min = 0xffffffffu; rk_max = 0u; rk_nan = 0u;
                        }
                        rk_b = b;
                        const unsigned k = mass_key_desc(m);
                        rk_min = k < rk_min ? k : rk_min; rk_max = k > rk_max ? k : rk_max; rk_nan |= (m != m);
                    }
                }
            }
            if (p.mass_range) {
                // one REDUX per value when the whole warp worked on the same batch row (always, for B = 1), else per thread
                const int b0 = __shfl_sync(0xffffffffu, rk_b, 0);
                if (__all_sync(0xffffffffu, rk_b == b0 || rk_b < 0)) {
                    const unsigned wmin = __reduce_min_sync(0xffffffffu, rk_min), wmax = __reduce_max_sync(0xffffffffu, rk_max);
                    const unsigned wnan = __reduce_or_sync(0xffffffffu, rk_nan);
                    const int bb = __reduce_max_sync(0xffffffffu, rk_b);
                    if (lane == 0 && bb >= 0) {
                        atomicMin(p.mass_range + 4 * bb, wmin); atomicMax(p.mass_range + 4 * bb + 1, wmax);
                        if (wnan) atomicOr(p.mass_range + 4 * bb + 2, 1u);
                    }
                } else if (rk_b >= 0) {
                    atomicMin(p.mass_range + 4 * rk_b, rk_min); atomicMax(p.mass_range + 4 * rk_b + 1, rk_max);
                    if (rk_nan) atomicOr(p.mass_range + 4 * rk_b + 2, 1u);
                }
            }
        }
    }
    if (p.mass) {       // only the epoch counter is left to advance (local; the flags were point-to-point)
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&p.state[1], 1u) == n_ctas - 1) {
            p.state[1] = 0u;
            __threadfence();
            ((volatile unsigned *)p.state)[0] = epoch;
        }
        return;
    }
    publish_epoch(p, parity, epoch, (unsigned)n_ctas);
}

template <int DT>
__global__ void __launch_bounds__(kStagedThreads, 1) shard_stats_staged_kernel(ShardParams p, StagedCfg c)
{
    shard_stats_staged_body<DT>(p, c, (int)blockIdx.x, (int)gridDim.x);
}

// Test harness for the one-kernel exchange on a single GPU: ALL ranks in ONE cooperative launch (blockIdx.y = rank), so
// that the CTAs that wait for each other's flags are guaranteed to be co-resident (separate launches that wait on one
// another are not: they may never run at the same time).  Same body, same code path as the per-rank kernel.
template <int DT>
__global__ void __launch_bounds__(kStagedThreads, 1) shard_stats_staged_allranks_kernel(const ShardParams *ps, StagedCfg c)
{
    const ShardParams p = ps[blockIdx.y];
    shard_stats_staged_body<DT>(p, c, (int)blockIdx.x, (int)gridDim.x);
}

__global__ void __launch_bounds__(256) shard_combine_kernel(const char *xlocal, unsigned *state, int64_t rows, int world,
                                                             int gpr, int rec_vecs, float fH, float *__restrict__ mass,
                                                             float *__restrict__ minmax)
{
    pdl_wait();
    pdl_trigger();
    const unsigned epoch = ((volatile unsigned *)state)[0];      // published by this rank's shard_stats_kernel
    const unsigned parity = epoch & 1u;
    if ((int)threadIdx.x < world) {
        const unsigned *flag = (const unsigned *)xlocal + parity * 8 + threadIdx.x;
        const long long t0 = clock64();
        while (ld_sys_u32(flag) != epoch) {                        // relaxed polling, one acquire fence afterwards
            if (clock64() - t0 > 40000000000LL) {                   // ~20 s: a peer never showed up; flag it, do not trap
                atomicExch(&state[2], 1u);
                break;
            }
        }
        fence_sys();
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    combine_token(xlocal, parity, world, gpr, rec_vecs, rows, t, fH, mass, minmax);
}

static int rec_vecs_for(int world) { return 1 + ((kGroups / world) + 3) / 4; }

} // namespace rtkv

using namespace rtkv;

static int g_shard_variant = 0;     // 0 automatic (staged when the shape allows), 1 always the generic register kernel
extern "C" int rtkv_debug_shard_variant(int variant)
{
    RTKV_CHECK_ARG(variant >= 0 && variant <= 3, "rtkv_debug_shard_variant: 0..3");
    g_shard_variant = variant;
    return RTKV_OK;
}

extern "C" int64_t rtkv_shard_exchange_bytes(int B, int S, int world)
{
    if (B <= 0 || S <= 0 || !(world == 1 || world == 2 || world == 4 || world == 8)) return -1;
    return (int64_t)kXHeader + (int64_t)2 * world * B * S * rec_vecs_for(world) * 16;
}

static int g_shard_l2_keep_mb = 0;     // MB of leading K/V rows kept in L2 by the statistics kernel (0: off)
static int g_shard_grid_cap = 0;    // testing aid: cap the persistent grid (several emulated ranks co-resident on one GPU)
extern "C" int rtkv_debug_shard_grid(int max_ctas)
{
    RTKV_CHECK_ARG(max_ctas >= 0, "rtkv_debug_shard_grid: negative");
    g_shard_grid_cap = max_ctas;
    return RTKV_OK;
}

// capture: when not NULL the call fills *capture_p / *capture_c with what it WOULD launch (staged kernel only) and launches
// nothing - the all-ranks test harness collects one parameter block per rank this way.
static int shard_stats_impl(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh, int64_t si,
                            int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                            const uint64_t *xbuf_ptrs, int rank, int world, uint32_t *state, int H_total, float *mass,
                            float *minmax, void *stream, ShardParams *capture_p = nullptr, StagedCfg *capture_c = nullptr,
                            int *capture_grid = nullptr, size_t *capture_smem = nullptr, uint32_t *mass_range = nullptr)
{
    RTKV_CHECK_ARG(A && K && V && xbuf_ptrs && state, "rtkv_shard_stats: null pointer");
    RTKV_CHECK_ARG(a_dtype == RTKV_F32, "rtkv_shard_stats: head sharding needs float32 probabilities "
                   "(half dtypes round the per-column head mean, which does not decompose over ranks)");
    RTKV_CHECK_ARG(dtype_ok(kv_dtype), "rtkv_shard_stats: bad K/V dtype %d", kv_dtype);
    RTKV_CHECK_ARG(B > 0 && H_local > 0 && S > 0 && P > 0 && group_size > 0 && n_local > 0, "rtkv_shard_stats: bad shape");
    RTKV_CHECK_ARG(world == 1 || world == 2 || world == 4 || world == 8, "rtkv_shard_stats: world must be 1, 2, 4 or 8");
    RTKV_CHECK_ARG(rank >= 0 && rank < world, "rtkv_shard_stats: bad rank");
    RTKV_CHECK_ARG(H_local == (kGroups / world) * group_size,
                   "rtkv_shard_stats: H_local=%d must be %d groups of group_size=%d heads", H_local, kGroups / world, group_size);
    const int epv = 16 / dtype_bytes(kv_dtype);
    if (n_local % epv != 0 || (uintptr_t)K % 16 || (uintptr_t)V % 16) {
        set_error("rtkv_shard_stats: K/V row slices must be 16-byte aligned vectors (n_local=%d)", n_local);
        return RTKV_EUNSUPPORTED;
    }
    ShardParams p;
    p.A = A; p.sb = sb; p.sh = sh; p.si = si; p.K = K; p.V = V;
    p.B = B; p.S = S; p.P = P; p.gs = group_size; p.ng = kGroups / world; p.n = n_local;
    p.rank = rank; p.world = world; p.rec_vecs = rec_vecs_for(world); p.dbg = g_shard_variant;
    // rows whose K/V slices are asked to stay in L2 for the quantise kernel (see the producer): RTKV_SHARD_L2_KEEP_MB of K+V
    // per batch row, default g_shard_l2_keep_mb
    {
        int keep_mb = g_shard_l2_keep_mb;
        if (const char *e = getenv("RTKV_SHARD_L2_KEEP_MB")) keep_mb = atoi(e);
        const double row_kv_bytes = 2.0 * n_local * dtype_bytes(kv_dtype);
        const double rows_keep = keep_mb > 0 ? keep_mb * 1048576.0 / (row_kv_bytes * B) : 0.0;
        p.kv_keep = rows_keep >= S ? S : (int)rows_keep;
    }
    for (int s = 0; s < 8; ++s) p.xbuf[s] = s < world ? xbuf_ptrs[s] : 0ull;
    for (int s = 0; s < world; ++s) RTKV_CHECK_ARG(p.xbuf[s] && p.xbuf[s] % 16 == 0, "rtkv_shard_stats: exchange buffer %d", s);
    p.state = state;
    p.mass = mass; p.minmax = minmax; p.fH = (float)H_total; p.mass_range = mass_range;
    const bool vec = ((uintptr_t)A % 16 == 0) && (sb * 4) % 16 == 0 && (sh * 4) % 16 == 0 && (si * 4) % 16 == 0;
    const int64_t rows = (int64_t)B * S;
    // staged path: every copy must be a 16-byte multiple at a 16-byte aligned address, P <= 128 columns, and the
    // prompt-column runs of consecutive rows must be adjacent (A is a [B,H,S,P] slice: si == P)
    StagedCfg c;
    c.kv_bytes = n_local * dtype_bytes(kv_dtype);
    c.a_bytes = P * 4;
    const int row_bytes = 2 * c.kv_bytes + H_local * c.a_bytes;
    const int budget = 200 * 1024;
    c.rb = budget / 2 / row_bytes;                              // at least two stages
    if (c.rb > kStagedConsumers) c.rb = kStagedConsumers;
    int want_stages = 4;
    if (const char *e = getenv("RTKV_SHARD_RB")) { const int v = atoi(e); if (v >= 1 && v <= c.rb) c.rb = v; }      // tuning knobs
    if (const char *e = getenv("RTKV_SHARD_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 8) want_stages = v; }
    const bool staged = g_shard_variant != 1 && vec && P % 4 == 0 && P <= 128 && si == P && c.kv_bytes % 16 == 0 && c.rb >= 1;
    if (staged) {
        c.stage_bytes = ((c.rb * row_bytes + 127) / 128) * 128;
        c.stages = budget / c.stage_bytes < want_stages ? budget / c.stage_bytes : want_stages;
        const size_t smem = kStagedHeader + (size_t)c.stages * c.stage_bytes;
        int dev = 0, sms = 0;
        RTKV_CUDA(cudaGetDevice(&dev));
        RTKV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t n_blocks = (int64_t)B * ((S + c.rb - 1) / c.rb);
        int grid = (int)(n_blocks < sms ? n_blocks : sms);
        if (g_shard_grid_cap > 0 && grid > g_shard_grid_cap) grid = g_shard_grid_cap;
        if (mass && grid > kMaxShardCtas) grid = kMaxShardCtas;
        if (capture_p) { *capture_p = p; *capture_c = c; *capture_grid = grid; *capture_smem = smem; return RTKV_OK; }
        void (*kern)(ShardParams, StagedCfg) = shard_stats_staged_kernel<RTKV_BF16>;
        if (kv_dtype == RTKV_F32) kern = shard_stats_staged_kernel<RTKV_F32>;
        else if (kv_dtype == RTKV_F16) kern = shard_stats_staged_kernel<RTKV_F16>;
        RTKV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStagedHeader + budget));
        RTKV_CUDA(launch_kernel(kern, dim3(grid), dim3(kStagedThreads), smem, (cudaStream_t)stream, 1, p, c));
        return RTKV_OK;
    }
    if (mass) {
        set_error("rtkv_shard_stats_combine: needs the staged kernel (A as a [B,H,S,P] slice with P %% 4 == 0, P <= 128, 16-byte "
                  "aligned rows); use rtkv_shard_stats + rtkv_shard_combine for this shape");
        return RTKV_EUNSUPPORTED;
    }
    const int64_t ctas = (rows + kShardWarps - 1) / kShardWarps;
    const int grid = (int)(ctas < 148 * 64 ? ctas : 148 * 64);
    void (*kern)(ShardParams) = nullptr;
    switch (kv_dtype) {
    case RTKV_F32: kern = vec ? shard_stats_generic_kernel<RTKV_F32, true> : shard_stats_generic_kernel<RTKV_F32, false>; break;
    case RTKV_F16: kern = vec ? shard_stats_generic_kernel<RTKV_F16, true> : shard_stats_generic_kernel<RTKV_F16, false>; break;
    default: kern = vec ? shard_stats_generic_kernel<RTKV_BF16, true> : shard_stats_generic_kernel<RTKV_BF16, false>; break;
    }
    RTKV_CUDA(launch_kernel(kern, dim3(grid), dim3(kShardWarps * 32), 0, (cudaStream_t)stream, 1, p));
    return RTKV_OK;
}

extern "C" int rtkv_shard_stats(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh, int64_t si,
                                int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                                const uint64_t *xbuf_ptrs, int rank, int world, uint32_t *state, void *stream)
{
    return shard_stats_impl(A, a_dtype, B, H_local, S, sb, sh, si, P, group_size, K, V, kv_dtype, n_local, xbuf_ptrs, rank, world,
                            state, 0, nullptr, nullptr, stream);
}

extern "C" int rtkv_shard_stats_combine(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh, int64_t si,
                                        int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                                        const uint64_t *xbuf_ptrs, int rank, int world, uint32_t *state, int H,
                                        float *mass, float *minmax, void *stream)
{
    RTKV_CHECK_ARG(mass && minmax && H > 0, "rtkv_shard_stats_combine: null output");
    RTKV_CHECK_ARG((uintptr_t)minmax % 16 == 0, "rtkv_shard_stats_combine: minmax must be 16-byte aligned");
    return shard_stats_impl(A, a_dtype, B, H_local, S, sb, sh, si, P, group_size, K, V, kv_dtype, n_local, xbuf_ptrs, rank, world,
                            state, H, mass, minmax, stream);
}

extern "C" int rtkv_shard_stats_combine_ranged(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh,
                                               int64_t si, int P, int group_size, const void *K, const void *V, int kv_dtype,
                                               int n_local, const uint64_t *xbuf_ptrs, int rank, int world, uint32_t *state, int H,
                                               float *mass, float *minmax, uint32_t *mass_range, void *stream)
{
    RTKV_CHECK_ARG(mass && minmax && mass_range && H > 0, "rtkv_shard_stats_combine_ranged: null output");
    RTKV_CHECK_ARG((uintptr_t)minmax % 16 == 0, "rtkv_shard_stats_combine_ranged: minmax must be 16-byte aligned");
    return shard_stats_impl(A, a_dtype, B, H_local, S, sb, sh, si, P, group_size, K, V, kv_dtype, n_local, xbuf_ptrs, rank, world,
                            state, H, mass, minmax, stream, nullptr, nullptr, nullptr, nullptr, mass_range);
}

extern "C" int rtkv_debug_shard_stats_combine_allranks(const void *const *A, int a_dtype, int B, int H_local, int S, int64_t sb,
                                                       int64_t sh, int64_t si, int P, int group_size, const void *const *K,
                                                       const void *const *V, int kv_dtype, int n_local, const uint64_t *xbuf_ptrs,
                                                       int world, uint32_t *const *state, int H, float *const *mass,
                                                       float *const *minmax, uint32_t *const *mass_range, void *params_scratch,
                                                       void *stream)
{
    RTKV_CHECK_ARG(A && K && V && xbuf_ptrs && state && mass && minmax && params_scratch, "rtkv_debug_shard_stats_combine_allranks: null pointer");
    RTKV_CHECK_ARG(world >= 1 && world <= 8, "rtkv_debug_shard_stats_combine_allranks: bad world");
    ShardParams ps[8];
    StagedCfg c;
    int grid = 0;
    size_t smem = 0;
    int dev = 0, sms = 0;
    RTKV_CUDA(cudaGetDevice(&dev));
    RTKV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int saved_cap = g_shard_grid_cap;
    g_shard_grid_cap = sms / world > 0 ? sms / world : 1;       // all ranks' CTAs must be co-resident: one CTA per SM
    int rc = RTKV_OK;
    for (int r = 0; r < world && !rc; ++r)
        rc = shard_stats_impl(A[r], a_dtype, B, H_local, S, sb, sh, si, P, group_size, K[r], V[r], kv_dtype, n_local, xbuf_ptrs, r,
                              world, state[r], H, mass[r], minmax[r], stream, &ps[r], &c, &grid, &smem, mass_range ? mass_range[r] : nullptr);
    g_shard_grid_cap = saved_cap;
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    RTKV_CUDA(cudaMemcpyAsync(params_scratch, ps, sizeof(ShardParams) * world, cudaMemcpyHostToDevice, st));
    RTKV_CUDA(cudaStreamSynchronize(st));                          // ps lives on this stack frame
    const ShardParams *dps = (const ShardParams *)params_scratch;
    void *args[2] = {(void *)&dps, (void *)&c};
    const void *kern = (const void *)shard_stats_staged_allranks_kernel<RTKV_BF16>;
    if (kv_dtype == RTKV_F32) kern = (const void *)shard_stats_staged_allranks_kernel<RTKV_F32>;
    else if (kv_dtype == RTKV_F16) kern = (const void *)shard_stats_staged_allranks_kernel<RTKV_F16>;
    RTKV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStagedHeader + 200 * 1024));
    RTKV_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid, world), dim3(kStagedThreads), args, smem, st));
    return RTKV_OK;
}

extern "C" int rtkv_shard_combine(const void *xbuf_local, uint32_t *state, int B, int S, int H, int world,
                                  float *mass, float *minmax, void *stream)
{
    RTKV_CHECK_ARG(xbuf_local && state && mass && minmax, "rtkv_shard_combine: null pointer");
    RTKV_CHECK_ARG(B > 0 && S > 0 && H > 0, "rtkv_shard_combine: bad shape");
    RTKV_CHECK_ARG(world == 1 || world == 2 || world == 4 || world == 8, "rtkv_shard_combine: world must be 1, 2, 4 or 8");
    RTKV_CHECK_ARG((uintptr_t)minmax % 16 == 0, "rtkv_shard_combine: minmax must be 16-byte aligned");
    const int64_t rows = (int64_t)B * S;
    RTKV_CUDA(launch_kernel(shard_combine_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1,
                            (const char *)xbuf_local, (unsigned *)state, rows, world, kGroups / world, rec_vecs_for(world),
                            (float)H, mass, minmax));
    return RTKV_OK;
}
