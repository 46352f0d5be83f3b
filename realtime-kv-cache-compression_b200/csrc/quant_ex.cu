// quant_ex.cu — extended scale granularities and outlier split for the quantise+gather stage.
//
// BASELINE.json north star, subsystem (2): "per-channel K and per-token V group-wise scale/zero
// computation and outlier detection and split", bit-packing with 128-bit accesses.  The reference
// itself only has one scale per whole [H*D] token row (src/compression/dynamic_quantization.py:183-184,
// quant_gather.cu); everything here is an extension whose semantics are defined by
// oracle/rtkv_oracle.c orc_fake_quant_ex (see include/rtkv.h rtkv_quant_options).
//
// Same streaming structure as K3: each kept element is read once (128-bit, L1::no_allocate) and
// written once; the data a scale depends on stays in registers between the statistics and the
// quantise pass:
//   token modes    one CTA per (kept row, tensor); a group of channels is a run of consecutive lanes,
//                  so group statistics are xor-shuffle segments (no shared memory, no barrier)
//   channel mode   one thread owns a 16-byte channel column of a group of <= 32 kept rows: min/max
//                  over tokens is thread-local, no cross-thread traffic at all
#include "quant_common.cuh"

namespace rtkv {

struct QXParams {
    QGParams g;
    int t;              // tensor handled by this launch (0 = X0, 1 = X1); -1: both, tensor = blockIdx.z (% 2)
    int group;          // token modes: channels per group (n for RTKV_Q_TOKEN)
    int lanes;          // token modes, vector kernel: 16-byte vectors per group (power of two <= 32), 0 = whole row
    int group1, lanes1; // the same for X1 when t == -1
    int tgroup;         // channel mode: kept rows per group
    unsigned tau_q8;    // 0 = no outlier split
    int32_t *ocount;    // [B,out_rows,2] or NULL
    // packed variants
    int64_t row_bytes;
    uint8_t *codes;
    float *qparams;
    int32_t *out_idx;
    float *out_val;
    int32_t *out_cnt;
    int cap;
};

// |x| * 2^(30-E) truncated to an integer, E = exponent field of the group's largest magnitude
__device__ __forceinline__ unsigned mag_fix(unsigned ab, unsigned E)
{
    const unsigned e = ab >> 23;
    const unsigned sh = E - e;
    if (e == 0 || sh >= 31) return 0;
    return (((ab & 0x7fffffu) | 0x800000u) << 7) >> sh;
}
__device__ __forceinline__ unsigned absbits(float f) { return __float_as_uint(f) & 0x7fffffffu; }

__device__ __forceinline__ unsigned block_max_u32(unsigned v, unsigned *sm)
{
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned r = sm[0];
#pragma unroll
    for (int w = 1; w < kQT / 32; ++w) r = max(r, sm[w]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *sm)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = 0;
#pragma unroll
    for (int w = 0; w < kQT / 32; ++w) r += sm[w];
    __syncthreads();
    return r;
}

// ---- token modes, vector path -------------------------------------------------------------------
// Issue-bound like K3 (ncu: K3 runs at 78 % issue utilisation), so instructions are what is budgeted here:
// segment statistics are single REDUX instructions on order-preserving integer images of the floats (one
// per slab and statistic, with the group's lane mask) instead of 5-step shuffle ladders, the outlier logic is
// compiled out when tau == 0 (TAU), the threshold is a 32-bit compare against a cached per-element mask, and
// the mean of a power-of-two group is a shift.
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// threshold as a 32-bit number: magnitudes are < 2^31, so 0xffffffff means "no outliers in this group"
__device__ __forceinline__ unsigned outlier_threshold(unsigned long long mean, unsigned tau_q8)
{
    const unsigned long long T = ((unsigned long long)tau_q8 * mean) >> 8;
    return T > 0xfffffffeull ? 0xffffffffu : (unsigned)T;
}

template <int DT, int NV, bool TAU>
__global__ void __launch_bounds__(kQT) quant_token_ex_kernel(QXParams q)
{
    constexpr int EPV = Vec16<DT>::kElems;
    __shared__ float red[16];
    __shared__ unsigned red_u[kQT / 32];
    __shared__ unsigned long long red_s[kQT / 32];
    __shared__ int s_count;
    const QGParams &p = q.g;
    const int r = blockIdx.x, b = blockIdx.y, t = q.t < 0 ? (int)blockIdx.z : q.t;
    const int group = (q.t < 0 && t) ? q.group1 : q.group;
    const int lanes = (q.t < 0 && t) ? q.lanes1 : q.lanes;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    uint4 *dst = (uint4 *)((char *)pick_y(p, t) + ((size_t)b * p.out_rows + r) * (size_t)p.n * Elem<DT>::kBytes);
    int32_t *oc = q.ocount ? q.ocount + ((size_t)b * p.out_rows + r) * 2 + t : nullptr;
    if (r >= cnt) {
        if (p.zero_pad)
            for (int v = threadIdx.x; v < nvec; v += kQT) stg_stream(dst + v, make_uint4(0, 0, 0, 0));
        if (oc && threadIdx.x == 0) *oc = 0;
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const uint4 *src = (const uint4 *)((const char *)pick_x(p, t) + ((size_t)b * p.S + tok) * (size_t)p.n * Elem<DT>::kBytes);
    uint4 raw[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = threadIdx.x + k * kQT;
        raw[k] = v < nvec ? ldg_stream(src + v) : make_uint4(0, 0, 0, 0);
    }
    if (oc && threadIdx.x == 0) s_count = 0;
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    const float qmax_r = qmax_rounded<DT>(bits);
    int n_out = 0;

    if (lanes == 0) {
        // ---- one group = the whole row: block-wide statistics
        int nan = 0;
        unsigned maxab = 0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (threadIdx.x + k * kQT < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); if (TAU) maxab = max(maxab, absbits(f[e])); }
            }
        }
        unsigned T = 0xffffffffu, E = 0;
        if (TAU) {
            maxab = block_max_u32(maxab, red_u);
            const int any_nan = __syncthreads_or(nan);
            E = maxab >> 23;
            if (!any_nan && E > 0) {
                unsigned long long sum = 0;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    if (threadIdx.x + k * kQT < nvec) {
                        float f[EPV];
                        Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) sum += mag_fix(absbits(f[e]), E);
                    }
                }
                sum = block_sum_u64(sum, red_s);
                T = outlier_threshold(sum / (unsigned long long)p.n, q.tau_q8);
            }
        }
        float mn = INFINITY, mx = -INFINITY;
        unsigned long long om = 0;                 // outlier bit per held element
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (threadIdx.x + k * kQT < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    if (TAU && mag_fix(absbits(f[e]), E) > T) { om |= 1ull << (k * EPV + e); continue; }
                    mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]);
                }
            }
        }
        block_minmax(mn, mx, nan, red);
        RowQuant<DT> rq;
        rq.init(mn, mx, bits);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * kQT;
            if (v < nvec) {
                float f[EPV], g[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) g[e] = f[e];
                if (rq.fast) rq.apply_fast(g);      // outliers lie outside [mn, mx]: their lanes are discarded below
                else {
#pragma unroll
                    for (int e = 0; e < EPV; ++e) g[e] = rq.slow(g[e]);
                }
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    if (!(TAU && ((om >> (k * EPV + e)) & 1))) f[e] = g[e];
                stg_stream(dst + v, Vec16<DT>::pack(f));
            }
        }
        if (TAU) n_out = __popcll(om);
    } else {
        // ---- a group is `lanes` (power of two) consecutive lanes of one 256-vector slab
        const unsigned lane = threadIdx.x & 31;
        const unsigned seg = lanes >= 32 ? 0xffffffffu : (((1u << lanes) - 1u) << (lane & ~(unsigned)(lanes - 1)));
        const int gshift = 31 - __clz(group);     // group = lanes * EPV is a power of two
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * kQT;
            float f[EPV];
            Vec16<DT>::unpack(raw[k], f);
            int nan = 0;
            unsigned maxab = 0;
#pragma unroll
            for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); if (TAU) maxab = max(maxab, absbits(f[e])); }
            nan = (__ballot_sync(0xffffffffu, nan) & seg) != 0;
            unsigned T = 0xffffffffu, E = 0;
            if (TAU) {
                E = __reduce_max_sync(seg, maxab) >> 23;
                unsigned long long sum = 0;
#pragma unroll
                for (int e = 0; e < EPV; ++e) sum += mag_fix(absbits(f[e]), E);
                // exact 64-bit segment sum from two 32-bit REDUX adds (sum < 2^34 per lane)
                const unsigned lo = __reduce_add_sync(seg, (unsigned)(sum & 0xfffffu));
                const unsigned hi = __reduce_add_sync(seg, (unsigned)(sum >> 20));
                sum = (unsigned long long)lo + ((unsigned long long)hi << 20);
                if (!nan && E > 0) T = outlier_threshold(sum >> gshift, q.tau_q8);
            }
            float mn = INFINITY, mx = -INFINITY;
            unsigned om = 0;
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                if (TAU && mag_fix(absbits(f[e]), E) > T) { om |= 1u << e; continue; }
                mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]);
            }
            mn = ord2f(__reduce_min_sync(seg, f2ord(mn)));
            mx = ord2f(__reduce_max_sync(seg, f2ord(mx)));
            if (nan) mn = mx = __int_as_float(0x7fc00000);
            if (v < nvec) {                        // uniform within a group (nvec is a multiple of lanes)
                RowQuant<DT> rq;
                rq.init(mn, mx, bits);
                float g[EPV];
#pragma unroll
                for (int e = 0; e < EPV; ++e) g[e] = f[e];
                if (rq.fast) rq.apply_fast(g);
                else {
#pragma unroll
                    for (int e = 0; e < EPV; ++e) g[e] = rq.slow(g[e]);
                }
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    if (!(TAU && ((om >> e) & 1))) f[e] = g[e];
                stg_stream(dst + v, Vec16<DT>::pack(f));
                if (TAU) n_out += __popc(om);
            }
        }
    }
    if (TAU && oc) {
        __syncthreads();                       // s_count = 0 is visible
        n_out = warp_sum(n_out);
        if ((threadIdx.x & 31) == 0 && n_out) atomicAdd(&s_count, n_out);
        __syncthreads();
        if (threadIdx.x == 0) *oc = s_count;
    } else if (oc && threadIdx.x == 0) {
        *oc = 0;
    }
}

// ---- token modes, any n / group / alignment: groups one after the other, block-wide passes ----
template <int DT>
__global__ void __launch_bounds__(kQT) quant_token_ex_generic_kernel(QXParams q)
{
    __shared__ float red[16];
    __shared__ unsigned red_u[kQT / 32];
    __shared__ unsigned long long red_s[kQT / 32];
    __shared__ int s_count;
    const QGParams &p = q.g;
    const int r = blockIdx.x, b = blockIdx.y, t = q.t;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int64_t dst = ((int64_t)b * p.out_rows + r) * p.n;
    int32_t *oc = q.ocount ? q.ocount + ((size_t)b * p.out_rows + r) * 2 + t : nullptr;
    if (r >= cnt) {
        if (p.zero_pad)
            for (int c = threadIdx.x; c < p.n; c += kQT) store_elem<DT>(pick_y(p, t), dst + c, 0.0f);
        if (oc && threadIdx.x == 0) *oc = 0;
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int64_t src = ((int64_t)b * p.S + tok) * p.n;
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    const float qmax_r = qmax_rounded<DT>(bits);
    if (threadIdx.x == 0) s_count = 0;
    int n_out = 0;
    for (int c0 = 0; c0 < p.n; c0 += q.group) {
        int nan = 0;
        unsigned maxab = 0;
        for (int c = c0 + threadIdx.x; c < c0 + q.group; c += kQT) {
            const float v = load_elem<DT>(pick_x(p, t), src + c);
            nan |= (v != v); maxab = max(maxab, absbits(v));
        }
        unsigned long long T = ~0ull;
        unsigned E = 0;
        bool use = false;
        if (q.tau_q8) {
            maxab = block_max_u32(maxab, red_u);
            const int any_nan = __syncthreads_or(nan);
            E = maxab >> 23;
            use = !any_nan && E > 0;
            unsigned long long sum = 0;
            for (int c = c0 + threadIdx.x; c < c0 + q.group; c += kQT)
                sum += mag_fix(absbits(load_elem<DT>(pick_x(p, t), src + c)), E);
            sum = block_sum_u64(sum, red_s);
            T = ((unsigned long long)q.tau_q8 * (sum / (unsigned long long)q.group)) >> 8;
        }
        float mn = INFINITY, mx = -INFINITY;
        for (int c = c0 + threadIdx.x; c < c0 + q.group; c += kQT) {
            const float v = load_elem<DT>(pick_x(p, t), src + c);
            if (use && mag_fix(absbits(v), E) > T) continue;
            mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
        block_minmax(mn, mx, nan, red);
        float scale, zp;
        quant_params<DT>(mn, mx, bits, scale, zp);
        for (int c = c0 + threadIdx.x; c < c0 + q.group; c += kQT) {
            const float v = load_elem<DT>(pick_x(p, t), src + c);
            if (use && mag_fix(absbits(v), E) > T) { ++n_out; store_elem<DT>(pick_y(p, t), dst + c, v); continue; }
            store_elem<DT>(pick_y(p, t), dst + c, dequant_code<DT>(quant_code<DT>(v, scale, zp, qmax_r), scale, zp));
        }
        __syncthreads();                       // red[] is reused by the next group
    }
    if (oc) {
        n_out = warp_sum(n_out);
        if ((threadIdx.x & 31) == 0 && n_out) atomicAdd(&s_count, n_out);
        __syncthreads();
        if (threadIdx.x == 0) *oc = s_count;
    }
}

// ---- channel mode -----------------------------------------------------------------------------
// One thread = one 16-byte channel column of a group of <= GT kept rows.  The tile is staged in shared
// memory with cp.async (global -> shared without passing through registers), so the kernel needs ~60
// registers and many small CTAs stay resident per SM: one CTA's loads overlap another's arithmetic and
// stores.  (The first version held the GT vectors in registers: 220 registers, 8 warps per SM, 18 % issue
// utilisation, 0.3 of the HBM peak - profiles/README.md.)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int DT, int GT, int CT>
__global__ void __launch_bounds__(CT) quant_channel_kernel(QXParams q)
{
    constexpr int EPV = Vec16<DT>::kElems;
    extern __shared__ __align__(16) unsigned char ch_smem[];
    uint4 *tile = (uint4 *)ch_smem + threadIdx.x;          // [GT][CT] vectors, this thread's column
    __shared__ int s_tok[GT], s_cls[GT];
    static_assert(CT >= GT, "CTA smaller than the token group");
    const QGParams &p = q.g;
    const int nslab = gridDim.z >> (q.t < 0 ? 1 : 0);
    const int r0 = blockIdx.x * q.tgroup, b = blockIdx.y, t = q.t < 0 ? (int)(blockIdx.z / nslab) : q.t;
    const int v = (blockIdx.z % nslab) * CT + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    if (threadIdx.x < GT) {
        const int r = r0 + threadIdx.x;
        int tok = -1, cls = 0;
        if (threadIdx.x < q.tgroup && r < cnt && r < p.out_rows) {
            tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
            cls = p.labels[(int64_t)b * p.S + tok];
        }
        s_tok[threadIdx.x] = tok; s_cls[threadIdx.x] = cls;
    }
    __syncthreads();
    if (v >= nvec) return;
    const size_t row_b = (size_t)p.n * Elem<DT>::kBytes;
    const char *xb = (const char *)pick_x(p, t) + (size_t)b * p.S * row_b;
    char *yb = (char *)pick_y(p, t) + (size_t)b * p.out_rows * row_b;
#pragma unroll
    for (int j = 0; j < GT; ++j)
        if (s_tok[j] >= 0) cp_async16(tile + j * CT, (const uint4 *)(xb + (size_t)s_tok[j] * row_b) + v);
    cp_async_wait_all();                                   // a thread only reads back what it copied itself
    float mn[EPV], mx[EPV];
    unsigned nanmask = 0;
#pragma unroll
    for (int e = 0; e < EPV; ++e) { mn[e] = INFINITY; mx[e] = -INFINITY; }
#pragma unroll 4
    for (int j = 0; j < GT; ++j) {
        if (s_tok[j] >= 0) {
            float f[EPV];
            Vec16<DT>::unpack(tile[j * CT], f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                nanmask |= (unsigned)(f[e] != f[e]) << e;
                mn[e] = fminf(mn[e], f[e]); mx[e] = fmaxf(mx[e], f[e]);
            }
        }
    }
    // scale / zero-point of every precision class for this thread's channels
    float sc0[EPV], zp0[EPV], sc1[EPV], zp1[EPV], sc2[EPV], zp2[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        if ((nanmask >> e) & 1) mn[e] = mx[e] = __int_as_float(0x7fc00000);
        quant_params<DT>(mn[e], mx[e], p.bits[0], sc0[e], zp0[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[1], sc1[e], zp1[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[2], sc2[e], zp2[e]);
    }
    const float qm0 = qmax_rounded<DT>(p.bits[0]), qm1 = qmax_rounded<DT>(p.bits[1]), qm2 = qmax_rounded<DT>(p.bits[2]);
#pragma unroll 2
    for (int j = 0; j < GT; ++j) {
        const int r = r0 + j;
        if (j >= q.tgroup || r >= p.out_rows) break;
        uint4 *dst = (uint4 *)(yb + (size_t)r * row_b) + v;
        if (s_tok[j] < 0) {
            if (p.zero_pad) stg_stream(dst, make_uint4(0, 0, 0, 0));
            continue;
        }
        const int cls = s_cls[j];
        const float qm = cls == 0 ? qm0 : (cls == 1 ? qm1 : qm2);
        float f[EPV];
        Vec16<DT>::unpack(tile[j * CT], f);
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            const float sc = cls == 0 ? sc0[e] : (cls == 1 ? sc1[e] : sc2[e]);
            const float z = cls == 0 ? zp0[e] : (cls == 1 ? zp1[e] : zp2[e]);
            f[e] = dequant_code<DT>(quant_code<DT>(f[e], sc, z, qm), sc, z);
        }
        stg_stream(dst, Vec16<DT>::pack(f));
    }
}

// Variant for long token groups: nothing is staged, the second pass re-reads the tile, which is still in L2 (a CTA touches
// GT rows x CT*16 bytes within a few microseconds): DRAM traffic stays algorithmic, the kernel needs neither the 32 KB of
// shared memory nor 4*GT registers per thread, and occupancy is bounded by ~50 registers only.
template <int DT, int GT, int CT>
__global__ void __launch_bounds__(CT) quant_channel_l2_kernel(QXParams q)
{
    constexpr int EPV = Vec16<DT>::kElems;
    constexpr int U = 8;                                    // loads in flight per thread
    __shared__ int s_tok[GT], s_cls[GT];
    static_assert(CT >= GT && GT % U == 0, "bad tile");
    const QGParams &p = q.g;
    const int nslab = gridDim.z >> (q.t < 0 ? 1 : 0);
    const int r0 = blockIdx.x * q.tgroup, b = blockIdx.y, t = q.t < 0 ? (int)(blockIdx.z / nslab) : q.t;
    const int v = (blockIdx.z % nslab) * CT + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    if (threadIdx.x < GT) {
        const int r = r0 + threadIdx.x;
        int tok = -1, cls = 0;
        if (threadIdx.x < q.tgroup && r < cnt && r < p.out_rows) {
            tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
            cls = p.labels[(int64_t)b * p.S + tok];
        }
        s_tok[threadIdx.x] = tok; s_cls[threadIdx.x] = cls;
    }
    __syncthreads();
    if (v >= nvec) return;
    const size_t row_b = (size_t)p.n * Elem<DT>::kBytes;
    const char *xb = (const char *)pick_x(p, t) + (size_t)b * p.S * row_b;
    char *yb = (char *)pick_y(p, t) + (size_t)b * p.out_rows * row_b;
    float mn[EPV], mx[EPV];
    unsigned nanmask = 0;
#pragma unroll
    for (int e = 0; e < EPV; ++e) { mn[e] = INFINITY; mx[e] = -INFINITY; }
#pragma unroll 1
    for (int j0 = 0; j0 < GT; j0 += U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (s_tok[j0 + u] >= 0) raw[u] = __ldg((const uint4 *)(xb + (size_t)s_tok[j0 + u] * row_b) + v);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (s_tok[j0 + u] >= 0) {
                float f[EPV];
                Vec16<DT>::unpack(raw[u], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    nanmask |= (unsigned)(f[e] != f[e]) << e;
                    mn[e] = fminf(mn[e], f[e]); mx[e] = fmaxf(mx[e], f[e]);
                }
            }
        }
    }
    float sc0[EPV], zp0[EPV], sc1[EPV], zp1[EPV], sc2[EPV], zp2[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        if ((nanmask >> e) & 1) mn[e] = mx[e] = __int_as_float(0x7fc00000);
        quant_params<DT>(mn[e], mx[e], p.bits[0], sc0[e], zp0[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[1], sc1[e], zp1[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[2], sc2[e], zp2[e]);
    }
    const float qm0 = qmax_rounded<DT>(p.bits[0]), qm1 = qmax_rounded<DT>(p.bits[1]), qm2 = qmax_rounded<DT>(p.bits[2]);
#pragma unroll 1
    for (int j0 = 0; j0 < GT; j0 += U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (s_tok[j0 + u] >= 0) raw[u] = ldg_stream((const uint4 *)(xb + (size_t)s_tok[j0 + u] * row_b) + v);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u, r = r0 + j;
            if (j >= q.tgroup || r >= p.out_rows) break;
            uint4 *dst = (uint4 *)(yb + (size_t)r * row_b) + v;
            if (s_tok[j] < 0) {
                if (p.zero_pad) stg_stream(dst, make_uint4(0, 0, 0, 0));
                continue;
            }
            const int cls = s_cls[j];
            const float qm = cls == 0 ? qm0 : (cls == 1 ? qm1 : qm2);
            float f[EPV];
            Vec16<DT>::unpack(raw[u], f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float sc = cls == 0 ? sc0[e] : (cls == 1 ? sc1[e] : sc2[e]);
                const float z = cls == 0 ? zp0[e] : (cls == 1 ? zp1[e] : zp2[e]);
                f[e] = dequant_code<DT>(quant_code<DT>(f[e], sc, z, qm), sc, z);
            }
            stg_stream(dst, Vec16<DT>::pack(f));
        }
    }
}

// channel mode, any n / alignment: one thread per channel, the group's rows are read twice
template <int DT>
__global__ void __launch_bounds__(kQT) quant_channel_generic_kernel(QXParams q)
{
    const QGParams &p = q.g;
    const int r0 = blockIdx.x * q.tgroup, b = blockIdx.y, t = q.t;
    const int c = blockIdx.z * kQT + threadIdx.x;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    if (c >= p.n) return;
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int j = 0; j < q.tgroup; ++j) {
        const int r = r0 + j;
        if (r >= cnt || r >= p.out_rows) break;
        const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
        const float v = load_elem<DT>(pick_x(p, t), ((int64_t)b * p.S + tok) * p.n + c);
        nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    if (nan) mn = mx = __int_as_float(0x7fc00000);
    for (int j = 0; j < q.tgroup; ++j) {
        const int r = r0 + j;
        if (r >= p.out_rows) break;
        const int64_t dst = ((int64_t)b * p.out_rows + r) * p.n + c;
        if (r >= cnt) {
            if (p.zero_pad) store_elem<DT>(pick_y(p, t), dst, 0.0f);
            continue;
        }
        const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
        const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
        float scale, zp;
        quant_params<DT>(mn, mx, bits, scale, zp);
        const float v = load_elem<DT>(pick_x(p, t), ((int64_t)b * p.S + tok) * p.n + c);
        store_elem<DT>(pick_y(p, t), dst, dequant_code<DT>(quant_code<DT>(v, scale, zp, qmax_rounded<DT>(bits)), scale, zp));
    }
}

// ---- packed rows, channel groups, vector path (no outlier split) ------------------------------------
// quant_token_ex_kernel's segment statistics with the packing epilogue of quant_pack_vec_kernel: a lane's
// 16-byte vector becomes EPV codes = one 1..16-byte store, the group's first lane writes (scale, zero-point).
template <int DT, int NV>
__global__ void __launch_bounds__(kQT) quant_pack_group_vec_kernel(QXParams q)
{
    constexpr int EPV = Vec16<DT>::kElems;
    const QGParams &p = q.g;
    const int r = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    const int G = p.n / q.group;
    const size_t orow = (size_t)b * p.out_rows + r;
    uint8_t *dst = q.codes + orow * (size_t)q.row_bytes;
    float *qp = q.qparams + orow * (size_t)G * 2;
    if (r >= cnt) {
        for (int64_t c = threadIdx.x; c < q.row_bytes / 16; c += kQT) ((uint4 *)dst)[c] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < 2 * G; i += kQT) qp[i] = 0.0f;
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const uint4 *src = (const uint4 *)((const char *)p.x[0] + ((size_t)b * p.S + tok) * (size_t)p.n * Elem<DT>::kBytes);
    uint4 raw[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = threadIdx.x + k * kQT;
        raw[k] = v < nvec ? ldg_stream(src + v) : make_uint4(0, 0, 0, 0);
    }
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    const int lanes = q.lanes;
    const unsigned lane = threadIdx.x & 31;
    const unsigned seg = lanes >= 32 ? 0xffffffffu : (((1u << lanes) - 1u) << (lane & ~(unsigned)(lanes - 1)));
    const int vbytes = bits > 8 ? 2 * EPV : EPV * bits / 8;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = threadIdx.x + k * kQT;
        float f[EPV];
        Vec16<DT>::unpack(raw[k], f);
        int nan = 0;
        float mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
        nan = (__ballot_sync(0xffffffffu, nan) & seg) != 0;
        mn = ord2f(__reduce_min_sync(seg, f2ord(mn)));
        mx = ord2f(__reduce_max_sync(seg, f2ord(mx)));
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        if (v < nvec) {
            RowQuant<DT> rq;
            rq.init(mn, mx, bits);
            if ((lane & (lanes - 1)) == 0) { const int gi = v / lanes; qp[2 * gi] = rq.scale; qp[2 * gi + 1] = rq.zp; }
            float cf[EPV];
            if (rq.fast) rq.codes_fast(f, cf);
            else {
#pragma unroll
                for (int e = 0; e < EPV; ++e) cf[e] = rq.code_slow(f[e]);
            }
            unsigned long long lo = 0, hi = 0;
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const unsigned long long c = cf[e] == cf[e] ? (unsigned long long)(unsigned)cf[e] : 0ull;
                const int sh = e * (bits > 8 ? 16 : bits);
                if (sh < 64) lo |= c << sh; else hi |= c << (sh - 64);
            }
            uint8_t *o = dst + (size_t)v * vbytes;
            if (vbytes == 16) *(uint4 *)o = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
            else if (vbytes == 8) *(unsigned long long *)o = lo;
            else if (vbytes == 4) *(unsigned *)o = (unsigned)lo;
            else if (vbytes == 2) *(uint16_t *)o = (uint16_t)lo;
            else *o = (uint8_t)lo;
        }
    }
    for (int64_t c = (int64_t)nvec * vbytes + threadIdx.x; c < q.row_bytes; c += kQT) dst[c] = 0;
}

template <int DT> static void (*pick_pack_group_kernel(int nv))(QXParams)
{
    if (nv <= 1) return quant_pack_group_vec_kernel<DT, 1>;
    if (nv <= 2) return quant_pack_group_vec_kernel<DT, 2>;
    if (nv <= 4) return quant_pack_group_vec_kernel<DT, 4>;
    if (nv <= 5) return quant_pack_group_vec_kernel<DT, 5>;
    return quant_pack_group_vec_kernel<DT, kMaxVec>;
}

// ---- packed rows, token modes -------------------------------------------------------------------
// One CTA per kept row, one warp per channel group (lanes stride the group's channels).  Phase 1 leaves
// every group's threshold / scale / zero-point / outlier count in shared memory, a scan turns the counts
// into list offsets (outliers are listed in channel order), phase 2 emits codes and the outlier list.
struct GroupRec { unsigned long long T; unsigned E; float scale, zp; int off; };

template <int DT>
__global__ void __launch_bounds__(kQT) quant_pack_ex_kernel(QXParams q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GroupRec *rec = (GroupRec *)smem_raw;
    __shared__ int s_total;
    const QGParams &p = q.g;
    const int r = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int G = p.n / q.group;
    const size_t orow = (size_t)b * p.out_rows + r;
    float *qp = q.qparams + orow * (size_t)G * 2;
    if (r >= cnt) {                                   // codes were zeroed by the host-side memset
        for (int i = threadIdx.x; i < 2 * G; i += kQT) qp[i] = 0.0f;
        if (q.out_cnt && threadIdx.x == 0) q.out_cnt[orow] = 0;
        return;
    }
    uint8_t *dst = q.codes + orow * (size_t)q.row_bytes;
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int64_t src = ((int64_t)b * p.S + tok) * p.n;
    const void *X = pick_x(p, q.t);
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    const float qmax_r = qmax_rounded<DT>(bits);

    for (int gi = warp; gi < G; gi += kQT / 32) {
        const int c0 = gi * q.group, c1 = c0 + q.group;
        int nan = 0;
        unsigned maxab = 0;
        for (int c = c0 + lane; c < c1; c += 32) {
            const float v = load_elem<DT>(X, src + c);
            nan |= (v != v); maxab = max(maxab, absbits(v));
        }
        nan = __any_sync(0xffffffffu, nan);
        unsigned long long T = ~0ull;
        unsigned E = 0;
        bool use = false;
        if (q.tau_q8) {
            maxab = __reduce_max_sync(0xffffffffu, maxab);
            E = maxab >> 23;
            use = !nan && E > 0;
            unsigned long long sum = 0;
            for (int c = c0 + lane; c < c1; c += 32) sum += mag_fix(absbits(load_elem<DT>(X, src + c)), E);
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            T = ((unsigned long long)q.tau_q8 * (sum / (unsigned long long)q.group)) >> 8;
        }
        float mn = INFINITY, mx = -INFINITY;
        int nout = 0;
        for (int c = c0 + lane; c < c1; c += 32) {
            const float v = load_elem<DT>(X, src + c);
            if (use && mag_fix(absbits(v), E) > T) { ++nout; continue; }
            mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
        mn = warp_min(mn); mx = warp_max(mx); nout = warp_sum(nout);
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        float scale, zp;
        quant_params<DT>(mn, mx, bits, scale, zp);
        if (lane == 0) {
            rec[gi].T = use ? T : ~0ull; rec[gi].E = E; rec[gi].scale = scale; rec[gi].zp = zp; rec[gi].off = nout;
            qp[2 * gi] = scale; qp[2 * gi + 1] = zp;
        }
    }
    __syncthreads();
    if (warp == 0) {                                   // exclusive scan of the per-group outlier counts
        const int per = (G + 31) / 32, g0 = lane * per;
        int run = 0;
        for (int i = g0; i < g0 + per && i < G; ++i) run += rec[i].off;
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        int base = incl - run;
        for (int i = g0; i < g0 + per && i < G; ++i) { const int c = rec[i].off; rec[i].off = base; base += c; }
        if (lane == 31) s_total = incl;
    }
    __syncthreads();
    const int cap = q.out_idx ? q.cap : 0;
    if (q.out_cnt && threadIdx.x == 0) q.out_cnt[orow] = s_total < cap ? s_total : cap;
    int32_t *oi = q.out_idx ? q.out_idx + orow * (size_t)cap : nullptr;
    float *ov = q.out_val ? q.out_val + orow * (size_t)cap : nullptr;
    const int per = bits > 8 ? 1 : 8 / bits;
    for (int gi = warp; gi < G; gi += kQT / 32) {
        const int c0 = gi * q.group, c1 = c0 + q.group;
        const unsigned long long T = rec[gi].T;
        const unsigned E = rec[gi].E;
        const float scale = rec[gi].scale, zp = rec[gi].zp;
        int base = rec[gi].off;
        for (int cb = c0; cb < c1; cb += 32) {
            const int c = cb + lane;
            const bool in = c < c1;
            const float v = in ? load_elem<DT>(X, src + c) : 0.0f;
            const bool is_out = in && T != ~0ull && mag_fix(absbits(v), E) > T;
            const unsigned ball = __ballot_sync(0xffffffffu, is_out);
            const int rank = base + __popc(ball & ((1u << lane) - 1u));
            base += __popc(ball);
            const bool split = is_out && rank < cap;
            unsigned code = 0;
            if (split) { oi[rank] = c; ov[rank] = v; }
            else if (in) { const float qf = quant_code<DT>(v, scale, zp, qmax_r); code = qf == qf ? (unsigned)qf : 0u; }
            if (bits > 8) {
                if (in) ((uint16_t *)dst)[c] = (uint16_t)code;
            } else {
                unsigned acc = code;
                for (int e = 1; e < per; ++e) acc |= __shfl_down_sync(0xffffffffu, code, e) << (e * bits);
                if (in && (lane % per) == 0) dst[c / per] = (uint8_t)acc;
            }
        }
    }
}

template <int DT>
__global__ void __launch_bounds__(kQT) unpack_dequant_ex_kernel(QXParams q)
{
    const QGParams &p = q.g;
    const int r = blockIdx.x, b = blockIdx.y;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const size_t orow = (size_t)b * p.out_rows + r;
    const int64_t dst = (int64_t)orow * p.n;
    void *Y = p.y[0];
    if (r >= cnt) {
        if (p.zero_pad) for (int c = threadIdx.x; c < p.n; c += kQT) store_elem<DT>(Y, dst + c, 0.0f);
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const uint8_t *src = q.codes + orow * (size_t)q.row_bytes;
    const int G = p.n / q.group;
    const float *qp = q.qparams + orow * (size_t)G * 2;
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        unsigned code;
        if (bits > 8) code = ((const uint16_t *)src)[c];
        else { const int per = 8 / bits; code = (src[c / per] >> ((c % per) * bits)) & ((1u << bits) - 1u); }
        const int gi = c / q.group;
        store_elem<DT>(Y, dst + c, dequant_code<DT>((float)code, qp[2 * gi], qp[2 * gi + 1]));
    }
    if (q.out_cnt) {
        __syncthreads();
        const int no = q.out_cnt[orow];
        for (int i = threadIdx.x; i < no; i += kQT)
            store_elem<DT>(Y, dst + q.out_idx[orow * (size_t)q.cap + i], q.out_val[orow * (size_t)q.cap + i]);
    }
}

// ---- packed rows, channel mode: one thread = 8 consecutive channels of a token group --------------
template <int DT, bool UNPACK>
__global__ void __launch_bounds__(kQT) pack_channel_kernel(QXParams q)
{
    const QGParams &p = q.g;
    const int g = blockIdx.x, r0 = g * q.tgroup, b = blockIdx.y;
    const int c8 = (blockIdx.z * kQT + threadIdx.x) * 8;
    if (c8 >= p.n) return;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int ngroups = (p.out_rows + q.tgroup - 1) / q.tgroup;
    float *qp = q.qparams + (((size_t)b * ngroups + g) * p.n + c8) * 2;
    float mn[8], mx[8];
    if (UNPACK) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { mn[e] = qp[2 * e]; mx[e] = qp[2 * e + 1]; }
    } else {
        unsigned nanmask = 0;
        bool any = false;
#pragma unroll
        for (int e = 0; e < 8; ++e) { mn[e] = INFINITY; mx[e] = -INFINITY; }
        for (int j = 0; j < q.tgroup; ++j) {
            const int r = r0 + j;
            if (r >= cnt || r >= p.out_rows) break;
            any = true;
            const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
            const int64_t src = ((int64_t)b * p.S + tok) * p.n + c8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float v = load_elem<DT>(p.x[0], src + e);
                nanmask |= (unsigned)(v != v) << e;
                mn[e] = fminf(mn[e], v); mx[e] = fmaxf(mx[e], v);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if ((nanmask >> e) & 1) mn[e] = mx[e] = __int_as_float(0x7fc00000);
            if (!any) mn[e] = mx[e] = 0.0f;
            qp[2 * e] = mn[e]; qp[2 * e + 1] = mx[e];
        }
    }
    for (int j = 0; j < q.tgroup; ++j) {
        const int r = r0 + j;
        if (r >= p.out_rows) break;
        const size_t orow = (size_t)b * p.out_rows + r;
        if (r >= cnt) {
            if (UNPACK && p.zero_pad)
                for (int e = 0; e < 8; ++e) store_elem<DT>(p.y[0], (int64_t)orow * p.n + c8 + e, 0.0f);
            continue;                                 // packing: codes were zeroed by the host-side memset
        }
        const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
        const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
        const float qmax_r = qmax_rounded<DT>(bits);
        uint8_t *row = q.codes + orow * (size_t)q.row_bytes;
        unsigned long long acc = 0;
        if (UNPACK && bits <= 8)
            for (int i = 0; i < bits; ++i) acc |= (unsigned long long)row[(size_t)c8 * bits / 8 + i] << (8 * i);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float scale, zp;
            quant_params<DT>(mn[e], mx[e], bits, scale, zp);
            if (UNPACK) {
                const unsigned code = bits > 8 ? ((const uint16_t *)row)[c8 + e]
                                               : (unsigned)((acc >> (e * bits)) & ((1ull << bits) - 1ull));
                store_elem<DT>(p.y[0], (int64_t)orow * p.n + c8 + e, dequant_code<DT>((float)code, scale, zp));
            } else {
                const float qf = quant_code<DT>(load_elem<DT>(p.x[0], ((int64_t)b * p.S + tok) * p.n + c8 + e), scale, zp, qmax_r);
                const unsigned code = qf == qf ? (unsigned)qf : 0u;
                if (bits > 8) ((uint16_t *)row)[c8 + e] = (uint16_t)code;
                else acc |= (unsigned long long)code << (e * bits);
            }
        }
        if (!UNPACK && bits <= 8)
            for (int i = 0; i < bits; ++i) row[(size_t)c8 * bits / 8 + i] = (uint8_t)(acc >> (8 * i));
    }
}

// Vector path of pack_channel_kernel<.., false>: one thread = a 16-byte channel column of a token group; pass 1 reads the
// column (min/max per channel, written as the group's parameter table), pass 2 re-reads it through L2 and emits each row's
// EPV codes as one 1..16-byte store.
template <int DT, int CT>
__global__ void __launch_bounds__(CT) pack_channel_vec_kernel(QXParams q)
{
    constexpr int EPV = Vec16<DT>::kElems;
    constexpr int GT = 32, U = 8;
    __shared__ int s_tok[GT], s_cls[GT];
    const QGParams &p = q.g;
    const int g = blockIdx.x, r0 = g * q.tgroup, b = blockIdx.y;
    const int v = blockIdx.z * CT + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    if (threadIdx.x < GT) {
        const int r = r0 + threadIdx.x;
        int tok = -1, cls = 0;
        if (threadIdx.x < q.tgroup && r < cnt && r < p.out_rows) {
            tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
            cls = p.labels[(int64_t)b * p.S + tok];
        }
        s_tok[threadIdx.x] = tok; s_cls[threadIdx.x] = cls;
    }
    __syncthreads();
    if (v >= nvec) return;
    const size_t row_b = (size_t)p.n * Elem<DT>::kBytes;
    const char *xb = (const char *)p.x[0] + (size_t)b * p.S * row_b;
    float mn[EPV], mx[EPV];
    unsigned nanmask = 0;
    bool any = false;
#pragma unroll
    for (int e = 0; e < EPV; ++e) { mn[e] = INFINITY; mx[e] = -INFINITY; }
#pragma unroll 1
    for (int j0 = 0; j0 < GT; j0 += U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (s_tok[j0 + u] >= 0) raw[u] = __ldg((const uint4 *)(xb + (size_t)s_tok[j0 + u] * row_b) + v);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (s_tok[j0 + u] >= 0) {
                any = true;
                float f[EPV];
                Vec16<DT>::unpack(raw[u], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    nanmask |= (unsigned)(f[e] != f[e]) << e;
                    mn[e] = fminf(mn[e], f[e]); mx[e] = fmaxf(mx[e], f[e]);
                }
            }
        }
    }
    const int ngroups = (p.out_rows + q.tgroup - 1) / q.tgroup;
    float *qp = q.qparams + (((size_t)b * ngroups + g) * p.n + (size_t)v * EPV) * 2;
    float sc0[EPV], zp0[EPV], sc1[EPV], zp1[EPV], sc2[EPV], zp2[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        if ((nanmask >> e) & 1) mn[e] = mx[e] = __int_as_float(0x7fc00000);
        if (!any) mn[e] = mx[e] = 0.0f;
        quant_params<DT>(mn[e], mx[e], p.bits[0], sc0[e], zp0[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[1], sc1[e], zp1[e]);
        quant_params<DT>(mn[e], mx[e], p.bits[2], sc2[e], zp2[e]);
    }
#pragma unroll
    for (int e = 0; e < EPV; e += 2)
        *(float4 *)(qp + 2 * e) = make_float4(mn[e], mx[e], mn[e + 1], mx[e + 1]);
    const float qm0 = qmax_rounded<DT>(p.bits[0]), qm1 = qmax_rounded<DT>(p.bits[1]), qm2 = qmax_rounded<DT>(p.bits[2]);
#pragma unroll 1
    for (int j0 = 0; j0 < GT; j0 += U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (s_tok[j0 + u] >= 0) raw[u] = ldg_stream((const uint4 *)(xb + (size_t)s_tok[j0 + u] * row_b) + v);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            if (s_tok[j] < 0) continue;                      // padding rows were zeroed by the host-side memset
            const int cls = s_cls[j];
            const int bits = cls == 0 ? p.bits[0] : (cls == 1 ? p.bits[1] : p.bits[2]);
            const float qm = cls == 0 ? qm0 : (cls == 1 ? qm1 : qm2);
            float f[EPV];
            Vec16<DT>::unpack(raw[u], f);
            unsigned long long lo = 0, hi = 0;
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float sc = cls == 0 ? sc0[e] : (cls == 1 ? sc1[e] : sc2[e]);
                const float z = cls == 0 ? zp0[e] : (cls == 1 ? zp1[e] : zp2[e]);
                const float qf = quant_code<DT>(f[e], sc, z, qm);
                const unsigned long long c = qf == qf ? (unsigned long long)(unsigned)qf : 0ull;
                const int sh = e * (bits > 8 ? 16 : bits);
                if (sh < 64) lo |= c << sh; else hi |= c << (sh - 64);
            }
            const int vbytes = bits > 8 ? 2 * EPV : EPV * bits / 8;
            uint8_t *o = q.codes + ((size_t)b * p.out_rows + r0 + j) * (size_t)q.row_bytes + (size_t)v * vbytes;
            if (vbytes == 16) *(uint4 *)o = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
            else if (vbytes == 8) *(unsigned long long *)o = lo;
            else if (vbytes == 4) *(unsigned *)o = (unsigned)lo;
            else if (vbytes == 2) *(uint16_t *)o = (uint16_t)lo;
            else *o = (uint8_t)lo;
        }
    }
}

__global__ void zero_counts_kernel(int32_t *oc, int64_t rows, int t)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) oc[i * 2 + t] = 0;
}

template <int DT> static bool vec_ok(const QGParams &p, int t)
{
    constexpr int EPV = Vec16<DT>::kElems;
    return p.n % EPV == 0 && (uintptr_t)p.x[t] % 16 == 0 && (uintptr_t)p.y[t] % 16 == 0;
}
// 16-byte vectors per group for the vector kernel; 0 = whole row; -1 = needs the generic kernel
template <int DT> static int group_lanes(const QGParams &p, int t, int group)
{
    constexpr int EPV = Vec16<DT>::kElems;
    if (!vec_ok<DT>(p, t) || p.n > kMaxVec * kQT * EPV) return -1;
    if (group == p.n) return 0;
    if (group % EPV) return -1;
    const int l = group / EPV;
    return (l <= 32 && (l & (l - 1)) == 0) ? l : -1;
}

// tensors [t0, t0+nt) share one launch (q.t = -1) when nt == 2
template <int DT> static int launch_token(QXParams q, int t0, int nt, cudaStream_t st)
{
    constexpr int EPV = Vec16<DT>::kElems;
    const QGParams &p = q.g;
    dim3 grid(p.out_rows, p.B, nt);
    const int nv = (p.n / EPV + kQT - 1) / kQT;
    void (*kern)(QXParams);
    if (q.tau_q8) {
        kern = quant_token_ex_kernel<DT, kMaxVec, true>;
        if (nv <= 1) kern = quant_token_ex_kernel<DT, 1, true>;
        else if (nv <= 2) kern = quant_token_ex_kernel<DT, 2, true>;
        else if (nv <= 4) kern = quant_token_ex_kernel<DT, 4, true>;
        else if (nv <= 5) kern = quant_token_ex_kernel<DT, 5, true>;
    } else {
        kern = quant_token_ex_kernel<DT, kMaxVec, false>;
        if (nv <= 1) kern = quant_token_ex_kernel<DT, 1, false>;
        else if (nv <= 2) kern = quant_token_ex_kernel<DT, 2, false>;
        else if (nv <= 4) kern = quant_token_ex_kernel<DT, 4, false>;
        else if (nv <= 5) kern = quant_token_ex_kernel<DT, 5, false>;
    }
    q.t = nt == 2 ? -1 : t0;
    RTKV_CUDA(launch_kernel(kern, grid, dim3(kQT), 0, st, 1, q));
    RTKV_LAUNCH_CHECK("quant_token_ex_kernel");
    return RTKV_OK;
}

static int g_channel_variant = 0;   // 0: L2 re-read kernel for 17..32-row groups, 1: cp.async-staged kernel

template <int DT> static int launch_channel(QXParams q, int t0, int nt, cudaStream_t st)
{
    constexpr int EPV = Vec16<DT>::kElems;
    const QGParams &p = q.g;
    const int ngroups = (p.out_rows + q.tgroup - 1) / q.tgroup;
    q.t = nt == 2 ? -1 : t0;
    if (q.tgroup <= 8) {
        dim3 grid(ngroups, p.B, (p.n / EPV + 127) / 128 * nt);
        RTKV_CUDA(launch_kernel(quant_channel_kernel<DT, 8, 128>, grid, dim3(128), 8 * 128 * 16, st, 1, q));
    } else if (q.tgroup <= 16) {
        dim3 grid(ngroups, p.B, (p.n / EPV + 127) / 128 * nt);
        RTKV_CUDA(launch_kernel(quant_channel_kernel<DT, 16, 128>, grid, dim3(128), 16 * 128 * 16, st, 1, q));
    } else if (g_channel_variant == 1) {
        dim3 grid(ngroups, p.B, (p.n / EPV + 63) / 64 * nt);
        RTKV_CUDA(launch_kernel(quant_channel_kernel<DT, 32, 64>, grid, dim3(64), 32 * 64 * 16, st, 1, q));
    } else {
        dim3 grid(ngroups, p.B, (p.n / EPV + 127) / 128 * nt);
        RTKV_CUDA(launch_kernel(quant_channel_l2_kernel<DT, 32, 128>, grid, dim3(128), 0, st, 1, q));
    }
    RTKV_LAUNCH_CHECK("quant_channel_kernel");
    return RTKV_OK;
}

template <int DT> static int launch_ex(QXParams q, const int *mode, const int *group, int nt, cudaStream_t st)
{
    const QGParams &p = q.g;
    int lanes[2] = {-1, -1};
    bool chan_vec[2] = {false, false};
    for (int t = 0; t < nt; ++t) {
        if (mode[t] == RTKV_Q_CHANNEL) chan_vec[t] = vec_ok<DT>(p, t);
        else lanes[t] = group_lanes<DT>(p, t, group[t]);
    }
    if (q.ocount)
        for (int t = 0; t < nt; ++t)
            if (mode[t] == RTKV_Q_CHANNEL) {
                const int64_t rows = (int64_t)p.B * p.out_rows;
                zero_counts_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(q.ocount, rows, t);
            }
    q.group = group[0]; q.lanes = lanes[0];
    q.group1 = nt == 2 ? group[1] : 0; q.lanes1 = nt == 2 ? lanes[1] : 0;
    if (nt == 2 && lanes[0] >= 0 && lanes[1] >= 0) return launch_token<DT>(q, 0, 2, st);
    if (nt == 2 && chan_vec[0] && chan_vec[1]) return launch_channel<DT>(q, 0, 2, st);
    for (int t = 0; t < nt; ++t) {
        q.group = group[t]; q.lanes = lanes[t]; q.t = t;
        int rc = RTKV_OK;
        if (mode[t] == RTKV_Q_CHANNEL) {
            if (chan_vec[t]) rc = launch_channel<DT>(q, t, 1, st);
            else {
                dim3 grid((p.out_rows + q.tgroup - 1) / q.tgroup, p.B, (p.n + kQT - 1) / kQT);
                quant_channel_generic_kernel<DT><<<grid, kQT, 0, st>>>(q);
                RTKV_LAUNCH_CHECK("quant_channel_generic_kernel");
            }
        } else if (lanes[t] >= 0) {
            rc = launch_token<DT>(q, t, 1, st);
        } else {
            quant_token_ex_generic_kernel<DT><<<dim3(p.out_rows, p.B), kQT, 0, st>>>(q);
            RTKV_LAUNCH_CHECK("quant_token_ex_generic_kernel");
        }
        if (rc) return rc;
    }
    return RTKV_OK;
}

static int check_options(const rtkv_quant_options *o, int n, int ntensors, const char *who)
{
    for (int t = 0; t < ntensors; ++t) {
        const int m = o->mode[t];
        if (m < RTKV_Q_TOKEN || m > RTKV_Q_CHANNEL) { set_error("%s: mode[%d]=%d", who, t, m); return RTKV_EINVAL; }
        if (m == RTKV_Q_TOKEN_GROUP && (o->group_size < 1 || n % o->group_size != 0)) {
            set_error("%s: group_size %d does not divide n=%d", who, o->group_size, n);
            return RTKV_EINVAL;
        }
        if (m == RTKV_Q_CHANNEL && (o->token_group < 1 || o->token_group > 32)) {
            set_error("%s: token_group %d outside 1..32", who, o->token_group);
            return RTKV_EINVAL;
        }
    }
    if (o->outlier_tau_q8 != 0 && o->outlier_tau_q8 < 256) {
        set_error("%s: outlier_tau_q8 %d: tau must be >= 1 (256)", who, o->outlier_tau_q8);
        return RTKV_EINVAL;
    }
    return RTKV_OK;
}

} // namespace rtkv

using namespace rtkv;

extern "C" int rtkv_debug_channel_variant(int v) { g_channel_variant = v; return RTKV_OK; }

extern "C" int rtkv_quant_groups(int n, const rtkv_quant_options *o, int tensor)
{
    if (n <= 0 || tensor < 0 || tensor > 1) return -1;
    if (!o || o->mode[tensor] == RTKV_Q_TOKEN) return 1;
    if (o->mode[tensor] == RTKV_Q_TOKEN_GROUP) return (o->group_size > 0 && n % o->group_size == 0) ? n / o->group_size : -1;
    return n;
}

extern "C" int rtkv_quant_gather_ex(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                    const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                    const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                    const rtkv_quant_options *opt, void *Y0, void *Y1, int32_t *outlier_counts,
                                    void *stream)
{
    RTKV_CHECK_ARG(X0 && Y0 && (!X1 || Y1), "rtkv_quant_gather_ex: null pointer");
    RTKV_CHECK_ARG(labels && bits, "rtkv_quant_gather_ex: labels and bits are required");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_quant_gather_ex: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_gather_ex: bad shape");
    RTKV_CHECK_ARG(B <= 65535, "rtkv_quant_gather_ex: B > 65535");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_gather_ex: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_quant_gather_ex: out_rows > S without sel_idx");
    int rc = rtkv_check_bits(bits, dtype, "rtkv_quant_gather_ex");
    if (rc) return rc;
    rtkv_quant_options o = {};
    if (opt) o = *opt;
    const int nt = X1 ? 2 : 1;
    rc = check_options(&o, n, nt, "rtkv_quant_gather_ex");
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool plain = o.outlier_tau_q8 == 0 && o.mode[0] == RTKV_Q_TOKEN && (nt == 1 || o.mode[1] == RTKV_Q_TOKEN);
    if (plain) {       // the reference's semantics: the K3 kernel itself
        if (outlier_counts) RTKV_CUDA(cudaMemsetAsync(outlier_counts, 0, (size_t)B * out_rows * 2 * sizeof(int32_t), st));
        return rtkv_quant_gather(X0, X1, dtype, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1, stream);
    }
    if (outlier_counts && nt == 1)
        RTKV_CUDA(cudaMemsetAsync(outlier_counts, 0, (size_t)B * out_rows * 2 * sizeof(int32_t), st));
    QXParams q = {};
    fill_params(q.g, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    q.tgroup = o.token_group; q.ocount = outlier_counts;
    q.tau_q8 = (unsigned)o.outlier_tau_q8;          // channel-mode tensors ignore it
    int group[2] = {n, n};
    for (int t = 0; t < nt; ++t)
        if (o.mode[t] == RTKV_Q_TOKEN_GROUP) group[t] = o.group_size;
    RTKV_DISPATCH(dtype, rc = launch_ex<DT>(q, o.mode, group, nt, st));
    return rc;
}

static int pack_ex_common(QXParams &q, const char *who, const void *X, void *Y, int dtype, int B, int S, int n,
                          const int32_t *sel_idx, const rtkv_row_stats *row_stats, const uint8_t *labels,
                          const int32_t *bits, int out_rows, int64_t row_bytes, const rtkv_quant_options *opt, int tensor,
                          const uint8_t *codes, const float *qparams, const int32_t *out_idx, const float *out_val,
                          const int32_t *out_cnt, int cap, int *mode)
{
    RTKV_CHECK_ARG(codes && qparams && labels && bits, "%s: null pointer", who);
    RTKV_CHECK_ARG(dtype_ok(dtype), "%s: bad dtype %d", who, dtype);
    RTKV_CHECK_ARG(B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "%s: bad shape", who);
    RTKV_CHECK_ARG(tensor == 0 || tensor == 1, "%s: tensor must be 0 or 1", who);
    RTKV_CHECK_ARG(!sel_idx || row_stats, "%s: sel_idx needs row_stats", who);
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "%s: out_rows > S without sel_idx", who);
    int rc = rtkv_check_bits(bits, dtype, who);
    if (rc) return rc;
    int mb = 0;
    for (int c = 0; c < 3; ++c) {
        const int bb = bits[c];
        if (!(bb == 1 || bb == 2 || bb == 4 || bb == 8 || (bb > 8 && bb <= 16))) {
            set_error("%s: packed rows need bits in {1,2,4,8,9..16}, got %d", who, bb);
            return RTKV_EUNSUPPORTED;
        }
        if (dtype == RTKV_BF16 && bb == 16) { set_error("%s: 16-bit codes are not representable for bf16 rows", who); return RTKV_EUNSUPPORTED; }
        mb = bb > mb ? bb : mb;
    }
    RTKV_CHECK_ARG(row_bytes >= rtkv_packed_row_bytes(n, mb), "%s: row_bytes too small", who);
    rtkv_quant_options o = {};
    if (opt) o = *opt;
    if (tensor == 1) o.mode[0] = o.mode[1];
    rc = check_options(&o, n, 1, who);
    if (rc) return rc;
    *mode = o.mode[0];
    q = QXParams{};
    fill_params(q.g, X, nullptr, B, S, n, sel_idx, row_stats, labels, bits, out_rows, 1, Y, nullptr);
    q.t = 0;
    q.group = *mode == RTKV_Q_TOKEN_GROUP ? o.group_size : n;
    q.tgroup = o.token_group;
    q.tau_q8 = *mode == RTKV_Q_CHANNEL ? 0u : (unsigned)o.outlier_tau_q8;
    q.row_bytes = row_bytes; q.codes = (uint8_t *)codes; q.qparams = (float *)qparams;
    q.out_idx = (int32_t *)out_idx; q.out_val = (float *)out_val; q.out_cnt = (int32_t *)out_cnt; q.cap = cap;
    if (*mode == RTKV_Q_CHANNEL) {
        if (n % 8 != 0) { set_error("%s: RTKV_Q_CHANNEL packing needs n %% 8 == 0", who); return RTKV_EUNSUPPORTED; }
    } else {
        if (*mode == RTKV_Q_TOKEN_GROUP && o.group_size % 8 != 0) {
            set_error("%s: packed group_size must be a multiple of 8", who);
            return RTKV_EUNSUPPORTED;
        }
        if (n / q.group > 1024) { set_error("%s: more than 1024 groups per row", who); return RTKV_EUNSUPPORTED; }
        if (q.tau_q8) RTKV_CHECK_ARG(out_idx && out_val && out_cnt && cap > 0, "%s: outlier split needs out_idx/out_val/out_cnt and a cap", who);
        else { q.out_idx = nullptr; q.out_val = nullptr; q.cap = 0; }
    }
    return RTKV_OK;
}

extern "C" int rtkv_quant_pack_ex(const void *X, int dtype, int B, int S, int n,
                                  const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                  const uint8_t *labels, const int32_t *bits, int out_rows, int64_t row_bytes,
                                  const rtkv_quant_options *opt, int tensor, uint8_t *codes, float *qparams,
                                  int32_t *out_idx, float *out_val, int32_t *out_cnt, int cap, void *stream)
{
    RTKV_CHECK_ARG(X, "rtkv_quant_pack_ex: null pointer");
    QXParams q;
    int mode = 0;
    int rc = pack_ex_common(q, "rtkv_quant_pack_ex", X, nullptr, dtype, B, S, n, sel_idx, row_stats, labels, bits, out_rows,
                            row_bytes, opt, tensor, codes, qparams, out_idx, out_val, out_cnt, cap, &mode);
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == RTKV_Q_TOKEN && !q.tau_q8) {          // the reference's scales: same layout as rtkv_quant_pack
        if (out_cnt) RTKV_CUDA(cudaMemsetAsync(out_cnt, 0, (size_t)B * out_rows * sizeof(int32_t), st));
        return rtkv_quant_pack(X, dtype, B, S, n, sel_idx, row_stats, labels, bits, out_rows, row_bytes, codes, qparams, stream);
    }
    const size_t codes_bytes = (size_t)B * out_rows * (size_t)row_bytes;
    if (mode == RTKV_Q_CHANNEL) {
        RTKV_CUDA(cudaMemsetAsync(codes, 0, codes_bytes, st));
        if (out_cnt) RTKV_CUDA(cudaMemsetAsync(out_cnt, 0, (size_t)B * out_rows * sizeof(int32_t), st));
        const int epv = 16 / dtype_bytes(dtype);
        int minb = bits[0] < bits[1] ? bits[0] : bits[1]; minb = minb < bits[2] ? minb : bits[2];
        const bool vec = n % epv == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)codes % 16 == 0 && (uintptr_t)qparams % 16 == 0 &&
                         row_bytes % 16 == 0 && (epv * minb) % 8 == 0;
        if (vec) {
            dim3 grid((out_rows + q.tgroup - 1) / q.tgroup, B, (n / epv + 127) / 128);
            RTKV_DISPATCH(dtype, (pack_channel_vec_kernel<DT, 128><<<grid, 128, 0, st>>>(q)));
        } else {
            dim3 grid((out_rows + q.tgroup - 1) / q.tgroup, B, (n / 8 + kQT - 1) / kQT);
            RTKV_DISPATCH(dtype, (pack_channel_kernel<DT, false><<<grid, kQT, 0, st>>>(q)));
        }
    } else {
        if (!q.tau_q8 && out_cnt) RTKV_CUDA(cudaMemsetAsync(out_cnt, 0, (size_t)B * out_rows * sizeof(int32_t), st));
        if (!q.tau_q8) q.out_cnt = nullptr;
        if (!q.tau_q8) {                          // vector kernel when the shapes allow it
            const int epv = 16 / dtype_bytes(dtype);
            int minb = bits[0] < bits[1] ? bits[0] : bits[1]; minb = minb < bits[2] ? minb : bits[2];
            const int l = q.group % epv == 0 ? q.group / epv : 0;
            const bool vec = n % epv == 0 && n <= kMaxVec * kQT * epv && (uintptr_t)X % 16 == 0 && (uintptr_t)codes % 16 == 0 &&
                             row_bytes % 16 == 0 && (epv * minb) % 8 == 0 && l >= 1 && l <= 32 && (l & (l - 1)) == 0;
            if (vec) {
                q.lanes = l;
                void (*kern)(QXParams) = nullptr;
                RTKV_DISPATCH(dtype, kern = pick_pack_group_kernel<DT>((n / epv + kQT - 1) / kQT));
                RTKV_CUDA(launch_kernel(kern, dim3(out_rows, B), dim3(kQT), 0, st, 1, q));
                RTKV_LAUNCH_CHECK("quant_pack_group_vec_kernel");
                return RTKV_OK;
            }
        }
        RTKV_CUDA(cudaMemsetAsync(codes, 0, codes_bytes, st));
        if (q.tau_q8) {                           // list slots past out_cnt are defined (zero), not left uninitialised
            RTKV_CUDA(cudaMemsetAsync(out_idx, 0, (size_t)B * out_rows * cap * sizeof(int32_t), st));
            RTKV_CUDA(cudaMemsetAsync(out_val, 0, (size_t)B * out_rows * cap * sizeof(float), st));
        }
        const size_t smem = (size_t)(n / q.group) * sizeof(GroupRec);
        RTKV_DISPATCH(dtype, (quant_pack_ex_kernel<DT><<<dim3(out_rows, B), kQT, smem, st>>>(q)));
    }
    RTKV_LAUNCH_CHECK("rtkv_quant_pack_ex");
    return RTKV_OK;
}

extern "C" int rtkv_unpack_dequant_ex(const uint8_t *codes, const float *qparams, const int32_t *out_idx,
                                      const float *out_val, const int32_t *out_cnt, int cap,
                                      int dtype, int B, int S, int n,
                                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                      const uint8_t *labels, const int32_t *bits, int out_rows, int64_t row_bytes,
                                      const rtkv_quant_options *opt, int tensor, void *Y, void *stream)
{
    RTKV_CHECK_ARG(Y, "rtkv_unpack_dequant_ex: null pointer");
    QXParams q;
    int mode = 0;
    int rc = pack_ex_common(q, "rtkv_unpack_dequant_ex", nullptr, Y, dtype, B, S, n, sel_idx, row_stats, labels, bits, out_rows,
                            row_bytes, opt, tensor, codes, qparams, out_idx, out_val, out_cnt, cap, &mode);
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == RTKV_Q_CHANNEL) {
        dim3 grid((out_rows + q.tgroup - 1) / q.tgroup, B, (n / 8 + kQT - 1) / kQT);
        RTKV_DISPATCH(dtype, (pack_channel_kernel<DT, true><<<grid, kQT, 0, st>>>(q)));
    } else {
        if (!q.tau_q8) q.out_cnt = nullptr;
        RTKV_DISPATCH(dtype, (unpack_dequant_ex_kernel<DT><<<dim3(out_rows, B), kQT, 0, st>>>(q)));
    }
    RTKV_LAUNCH_CHECK("rtkv_unpack_dequant_ex");
    return RTKV_OK;
}
