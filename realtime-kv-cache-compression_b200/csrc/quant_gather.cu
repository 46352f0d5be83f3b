// quant_gather.cu — K3: per-token asymmetric min/max fake-quantisation fused
// with the compacting gather of the kept K/V rows.
//
// Reference semantics (paths relative to the reference repo):
//   src/compression/dynamic_quantization.py:62-126   scale/zero-point, round, clamp, dequantise
//   src/compression/dynamic_quantization.py:128-196  bits picked per token by its label
//   src/compression/selective_propagation.py:213-232 kept rows, ascending index, zero padded
// The reference quantises every token and then drops the unselected ones; here
// only the kept rows are read (each exactly once) and written (exactly once).
//
// HBM-bound streaming kernel: 2 * rho * B*S * n elements in, the same out.
// Fast path: one CTA of 256 threads per (output row, tensor); the row lives in
// registers (up to kMaxVec 16-byte vectors per thread), so the min/max pass and
// the quantise pass cost no extra memory traffic.  Loads/stores are 128-bit,
// coalesced, streaming (L1::no_allocate).
#include "quant_common.cuh"

namespace rtkv {

__device__ int g_bf16_packed = 1;     // rtkv_debug_bf16_packed(): 0 = element-wise bf16 code only (A/B testing)

// NV = 16-byte vectors per thread (n <= NV * kQT * elems_per_vec, n % elems_per_vec == 0)
template <int DT, int NV>
__global__ void __launch_bounds__(kQT) quant_gather_vec_kernel(QGParams p)
{
    constexpr int EPV = Vec16<DT>::kElems;
    __shared__ float red[16];
    const int r = blockIdx.x + p.row_begin, b = blockIdx.y + p.b_begin, t = blockIdx.z;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    uint4 *dst = (uint4 *)((char *)pick_y(p, t) + ((size_t)b * p.out_rows + r) * (size_t)p.n * Elem<DT>::kBytes);
    if (r >= cnt) {
        if (p.zero_pad)
            for (int v = threadIdx.x; v < nvec; v += kQT) stg_stream(dst + v, make_uint4(0, 0, 0, 0));
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
    if (!p.quantize) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * kQT;
            if (v < nvec) stg_stream(dst + v, raw[k]);
        }
        return;
    }
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    if constexpr (DT == RTKV_BF16) {            // packed min/max, NaN propagating: one instruction per element
        __nv_bfloat162 mn2 = __float2bfloat162_rn(INFINITY), mx2 = __float2bfloat162_rn(-INFINITY);
#pragma unroll
        for (int k = 0; k < NV; ++k)
            if (threadIdx.x + k * kQT < nvec) bf16x8_minmax(raw[k], mn2, mx2);
        const float a = __low2float(mn2), c = __high2float(mn2), d = __low2float(mx2), e = __high2float(mx2);
        nan = (a != a) | (c != c) | (d != d) | (e != e);
        mn = fminf(a, c); mx = fmaxf(d, e);
    } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (threadIdx.x + k * kQT < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
            }
        }
    }
    if (p.ext_minmax) {         // row min/max reduced over all ranks' head shards
        ext_row_minmax(p, b, r, tok, t, mn, mx);
    } else {
        block_minmax(mn, mx, nan, red);
    }
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    RowQuant<DT> rq;
    rq.init(mn, mx, bits);
    if constexpr (DT == RTKV_BF16) {
        Bf16RowPacked pq;
        pq.init(mn, mx, rq.scale, rq.zp, rq.qmax_r);
        if (pq.ok && g_bf16_packed) {           // uniform over the CTA (row-level)
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int v = threadIdx.x + k * kQT;
                if (v < nvec) stg_stream(dst + v, pq.apply(raw[k]));
            }
            return;
        }
    }
    if (rq.fast) {                              // uniform over the CTA (row-level)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * kQT;
            if (v < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
                rq.apply_fast(f);
                stg_stream(dst + v, Vec16<DT>::pack(f));
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * kQT;
            if (v < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) f[e] = rq.slow(f[e]);
                stg_stream(dst + v, Vec16<DT>::pack(f));
            }
        }
    }
}

// ---- KV-head sharding, fused exchange ------------------------------------------------------
// Same as quant_gather_vec_kernel, but the row's scale spans the head slices of ALL ranks
// (dynamic_quantization.py:183-184 quantises the whole [H*D] row).  Instead of a separate min/max
// pass + collective + second read, the CTA keeps its slice of the row in registers, pushes its local
// (min, -max) into every rank's mailbox over NVLink (peer stores into symmetric memory) and polls its
// own mailbox until all R partials of this row have arrived.  Mailbox words are self-validating
// {value, epoch} pairs (8-byte granules, NCCL-LL style), so no fence or flag reset is needed; the
// epoch lives in device memory and is bumped by the last CTA, which makes the kernel graph-replayable.
// All ranks process the same rows in the same order, so a CTA only ever waits for a peer CTA that has
// already been scheduled or will be without needing this CTA's resources.
__device__ __forceinline__ void st_volatile_v4(void *p, uint4 v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// min / max / any-NaN over a CTA of NW warps
template <int NW>
__device__ __forceinline__ void block_minmax_n(float &mn, float &mx, int nan, float *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { red[warp] = mn; red[8 + warp] = mx; }
    const int any_nan = __syncthreads_or(nan);
    float a = red[0], z = red[8];
#pragma unroll
    for (int w = 1; w < NW; ++w) { a = fminf(a, red[w]); z = fmaxf(z, red[8 + w]); }
    mn = a; mx = z;
    if (any_nan) mn = mx = __int_as_float(0x7fc00000);
}

// THREADS per CTA is small (64/128) on purpose: a CTA spends most of its life waiting for the peers'
// partials (~NVLink round trip), so what hides that latency is the number of ROWS in flight per SM
// (32 resident 64-thread CTAs), not the width of a CTA.
template <int DT, int NV, int THREADS>
__global__ void __launch_bounds__(THREADS) quant_gather_p2p_kernel(QGParams p)
{
    constexpr int EPV = Vec16<DT>::kElems;
    __shared__ float red[16];
    __shared__ float s_mm[2];
    const int r = blockIdx.x, b = blockIdx.y, t = blockIdx.z;
    pdl_wait();
    pdl_trigger();
    const unsigned epoch = ((volatile unsigned *)p.p2p_state)[0] + 1u;
    const int cnt = p.row_stats[b].n_selected;
    const int nvec = p.n / EPV;
    uint4 *dst = (uint4 *)((char *)pick_y(p, t) + ((size_t)b * p.out_rows + r) * (size_t)p.n * Elem<DT>::kBytes);
    if (r < cnt) {
        const int tok = p.sel_idx[(int64_t)b * p.S + r];
        const uint4 *src = (const uint4 *)((const char *)pick_x(p, t) + ((size_t)b * p.S + tok) * (size_t)p.n * Elem<DT>::kBytes);
        uint4 raw[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * THREADS;
            raw[k] = v < nvec ? ldg_stream(src + v) : make_uint4(0, 0, 0, 0);
        }
        float mn = INFINITY, mx = -INFINITY;
        int nan = 0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (threadIdx.x + k * THREADS < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
            }
        }
        block_minmax_n<THREADS / 32>(mn, mx, nan, red);
        // ---- exchange: slot[(parity, tensor, row, source rank)] in every rank's mailbox ----
        const size_t rows_total = (size_t)p.B * p.out_rows;
        const size_t slot0 = ((((size_t)(epoch & 1u) * 2 + t) * rows_total + (size_t)b * p.out_rows + r) * p.world);
        if ((int)threadIdx.x < p.world) {
            const int s = threadIdx.x;
            uint4 mine = make_uint4(__float_as_uint(mn), epoch, __float_as_uint(-mx), epoch);
            st_volatile_v4((uint4 *)p.mbox[s] + slot0 + p.rank, mine);                  // into rank s's mailbox
            const uint4 *in = (const uint4 *)p.mbox[p.rank] + slot0 + s;               // from rank s, in my mailbox
            uint4 got;
            const long long t0 = clock64();
            bool timed_out = false;
            do {
                got = ld_volatile_v4(in);
                if (clock64() - t0 > 40000000000LL) { timed_out = true; break; }        // ~20 s: a peer never showed up
            } while (got.y != epoch || got.w != epoch);
            if (timed_out) {        // never trap (that would kill the CUDA context): flag it and go on with the local values
                atomicExch(&p.p2p_state[2], 1u);
                got = make_uint4(__float_as_uint(mn), epoch, __float_as_uint(-mx), epoch);
            }
            float pmn = __uint_as_float(got.x), pnmx = __uint_as_float(got.z);
            int pnan = (pmn != pmn);
            const unsigned act = (1u << p.world) - 1u;
            for (int o = 4; o; o >>= 1) {                                               // world <= 8
                // a partner lane outside the mask (world < 8) must not be read: shuffle from the own lane instead
                const int src = (s ^ o) < p.world ? (s ^ o) : s;
                const float omn = __shfl_sync(act, pmn, src), onm = __shfl_sync(act, pnmx, src);
                const int on = __shfl_sync(act, pnan, src);
                pmn = fminf(pmn, omn); pnmx = fminf(pnmx, onm); pnan |= on;
            }
            if (s == 0) {
                s_mm[0] = pnan ? __int_as_float(0x7fc00000) : pmn;
                s_mm[1] = pnan ? __int_as_float(0x7fc00000) : -pnmx;
            }
        }
        __syncthreads();
        mn = s_mm[0]; mx = s_mm[1];
        const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
        float scale, zp;
        quant_params<DT>(mn, mx, bits, scale, zp);
        const float qmax_r = qmax_rounded<DT>(bits);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int v = threadIdx.x + k * THREADS;
            if (v < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) f[e] = dequant_code<DT>(quant_code<DT>(f[e], scale, zp, qmax_r), scale, zp);
                stg_stream(dst + v, Vec16<DT>::pack(f));
            }
        }
    } else if (p.zero_pad) {
        for (int v = threadIdx.x; v < nvec; v += THREADS) stg_stream(dst + v, make_uint4(0, 0, 0, 0));
    }
    // last CTA of the grid publishes the epoch for the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&p.p2p_state[1], 1u) == (unsigned)(p.total_ctas - 1)) {
            p.p2p_state[1] = 0u;
            __threadfence();
            ((volatile unsigned *)p.p2p_state)[0] = epoch;
        }
    }
}

// ---- KV-head sharding: narrow row slices with externally reduced min/max ------------------------
// A rank's slice of a row is short (512 floats = 128 vectors when 8 ranks share a 4096-wide row) and its min/max
// comes from the exchange (rtkv_shard_combine), so no reduction over the row is needed: 2^tpr_shift consecutive
// threads own one (output row, tensor) item with exactly 4 vectors each, a 256-thread CTA processes
// 256 >> tpr_shift items and never synchronises.  Same arithmetic (RowQuant) as quant_gather_vec_kernel.
template <int DT>
__global__ void __launch_bounds__(kQT) quant_gather_flat_kernel(QGParams p, int tpr_shift, int ntensors)
{
    constexpr int EPV = Vec16<DT>::kElems;
    constexpr int NV = 4;
    pdl_wait();
    pdl_trigger();
    const int tpr = 1 << tpr_shift;
    const long long item = ((long long)blockIdx.x * kQT + threadIdx.x) >> tpr_shift;
    const int l = threadIdx.x & (tpr - 1);
    if (item >= (long long)p.B * p.out_rows * ntensors) return;
    const int t = (int)(item % ntensors);
    const long long rr = item / ntensors;
    const int r = (int)(rr % p.out_rows), b = (int)(rr / p.out_rows);
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    uint4 *dst = (uint4 *)((char *)pick_y(p, t) + ((size_t)b * p.out_rows + r) * (size_t)p.n * Elem<DT>::kBytes);
    if (r >= cnt) {
        if (p.zero_pad) {
#pragma unroll
            for (int k = 0; k < NV; ++k) stg_stream(dst + l + k * tpr, make_uint4(0, 0, 0, 0));
        }
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const uint4 *src = (const uint4 *)((const char *)pick_x(p, t) + ((size_t)b * p.S + tok) * (size_t)p.n * Elem<DT>::kBytes);
    uint4 raw[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) raw[k] = ldg_stream(src + l + k * tpr);
    float mn, mx;
    ext_row_minmax(p, b, r, tok, t, mn, mx);
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    RowQuant<DT> rq;
    rq.init(mn, mx, bits);
    if constexpr (DT == RTKV_BF16) {
        Bf16RowPacked pq;
        pq.init(mn, mx, rq.scale, rq.zp, rq.qmax_r);
        if (pq.ok && g_bf16_packed) {
#pragma unroll
            for (int k = 0; k < NV; ++k) stg_stream(dst + l + k * tpr, pq.apply(raw[k]));
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float f[EPV];
        Vec16<DT>::unpack(raw[k], f);
        if (rq.fast) rq.apply_fast(f);
        else {
#pragma unroll
            for (int e = 0; e < EPV; ++e) f[e] = rq.slow(f[e]);
        }
        stg_stream(dst + l + k * tpr, Vec16<DT>::pack(f));
    }
}

// Any n / alignment: scalar accesses, the row is read twice (second read hits L1/L2).
template <int DT>
__global__ void __launch_bounds__(kQT) quant_gather_generic_kernel(QGParams p)
{
    __shared__ float red[16];
    const int r = blockIdx.x + p.row_begin, b = blockIdx.y + p.b_begin, t = blockIdx.z;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int64_t dst = ((int64_t)b * p.out_rows + r) * p.n;
    if (r >= cnt) {
        if (p.zero_pad)
            for (int c = threadIdx.x; c < p.n; c += kQT) store_elem<DT>(pick_y(p, t), dst + c, 0.0f);
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int64_t src = ((int64_t)b * p.S + tok) * p.n;
    if (!p.quantize) {
        for (int c = threadIdx.x; c < p.n; c += kQT) store_elem<DT>(pick_y(p, t), dst + c, load_elem<DT>(pick_x(p, t), src + c));
        return;
    }
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        float v = load_elem<DT>(pick_x(p, t), src + c);
        nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    if (p.ext_minmax) {
        ext_row_minmax(p, b, r, tok, t, mn, mx);
    } else {
        block_minmax(mn, mx, nan, red);
    }
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    float scale, zp;
    quant_params<DT>(mn, mx, bits, scale, zp);
    const float qmax_r = qmax_rounded<DT>(bits);
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        float v = load_elem<DT>(pick_x(p, t), src + c);
        store_elem<DT>(pick_y(p, t), dst + c, dequant_code<DT>(quant_code<DT>(v, scale, zp, qmax_r), scale, zp));
    }
}

// Local (this rank's head shard) min/max of every kept row: out[b,r,:] = (min0, max0, min1, max1).
// Rows past n_selected get (+inf, -inf) so that a MIN/MAX reduction over ranks ignores them.
template <int DT>
__global__ void __launch_bounds__(kQT) row_minmax_kernel(QGParams p, float *out)
{
    __shared__ float red[16];
    const int r = blockIdx.x, b = blockIdx.y, t = blockIdx.z;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    float *o = out + ((size_t)b * p.out_rows + r) * 4 + 2 * t;
    if (r >= cnt) {
        if (threadIdx.x == 0) { o[0] = INFINITY; o[1] = p.neg_max ? INFINITY : -INFINITY; }
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int64_t src = ((int64_t)b * p.S + tok) * p.n;
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        const float v = load_elem<DT>(pick_x(p, t), src + c);
        nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    block_minmax(mn, mx, nan, red);
    if (threadIdx.x == 0) { o[0] = mn; o[1] = p.neg_max ? -mx : mx; }
}

// Vector path of row_minmax: same 128-bit streaming loads as the quantise kernel.
template <int DT>
__global__ void __launch_bounds__(kQT) row_minmax_vec_kernel(QGParams p, float *out)
{
    constexpr int EPV = Vec16<DT>::kElems;
    __shared__ float red[16];
    const int r = blockIdx.x, b = blockIdx.y, t = blockIdx.z;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    float *o = out + ((size_t)b * p.out_rows + r) * 4 + 2 * t;
    if (r >= cnt) {
        if (threadIdx.x == 0) { o[0] = INFINITY; o[1] = p.neg_max ? INFINITY : -INFINITY; }
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int nvec = p.n / EPV;
    const uint4 *src = (const uint4 *)((const char *)pick_x(p, t) + ((size_t)b * p.S + tok) * (size_t)p.n * Elem<DT>::kBytes);
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int v0 = threadIdx.x; v0 < nvec; v0 += 4 * kQT) {
        uint4 raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (v0 + k * kQT < nvec) raw[k] = ldg_stream(src + v0 + k * kQT);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (v0 + k * kQT < nvec) {
                float f[EPV];
                Vec16<DT>::unpack(raw[k], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
            }
        }
    }
    block_minmax(mn, mx, nan, red);
    if (threadIdx.x == 0) { o[0] = mn; o[1] = p.neg_max ? -mx : mx; }
}

template <int DT> static int launch_qg(const QGParams &p, int ntensors, cudaStream_t st, int nrows, int nb)
{
    constexpr int EPV = Vec16<DT>::kElems;
    dim3 grid(nrows, nb, ntensors);
    bool vec = p.n % EPV == 0 && p.n <= kMaxVec * kQT * EPV;
    for (int t = 0; t < ntensors; ++t)
        vec = vec && ((uintptr_t)p.x[t] % 16 == 0) && ((uintptr_t)p.y[t] % 16 == 0);
    if (vec) {
        const int nv = (p.n / EPV + kQT - 1) / kQT;
        void (*kern)(QGParams) = quant_gather_vec_kernel<DT, kMaxVec>;
        if (nv <= 1) kern = quant_gather_vec_kernel<DT, 1>;
        else if (nv <= 2) kern = quant_gather_vec_kernel<DT, 2>;
        else if (nv <= 3) kern = quant_gather_vec_kernel<DT, 3>;
        else if (nv <= 4) kern = quant_gather_vec_kernel<DT, 4>;
        else if (nv <= 5) kern = quant_gather_vec_kernel<DT, 5>;   // 13B rows: 5120 fp32
        else if (nv <= 6) kern = quant_gather_vec_kernel<DT, 6>;
        RTKV_CUDA(launch_kernel(kern, grid, dim3(kQT), 0, st, 1, p));
    } else {
        quant_gather_generic_kernel<DT><<<grid, kQT, 0, st>>>(p);
    }
    RTKV_LAUNCH_CHECK("quant_gather_kernel");
    return RTKV_OK;
}

// ---- whole-tensor parameters (component API: get_quantization_params / quantize_tensor) ----
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// scratch: [0] min (ordered uint), [1] max, [2] nan flag, [3] ticket
template <int DT>
__global__ void __launch_bounds__(kQT) tensor_minmax_kernel(const void *x, int64_t numel, int bits,
                                                            unsigned *scratch, float *qparams)
{
    __shared__ float red[16];
    __shared__ int last;
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int64_t i = (int64_t)blockIdx.x * kQT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * kQT) {
        float v = load_elem<DT>(x, i);
        nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    block_minmax(mn, mx, nan, red);
    if (threadIdx.x == 0) {
        if (mn != mn) atomicOr(&scratch[2], 1u);
        else { atomicMin(&scratch[0], f2ord(mn)); atomicMax(&scratch[1], f2ord(mx)); }
        __threadfence();
        last = atomicAdd(&scratch[3], 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            float a = ord2f(((volatile unsigned *)scratch)[0]), z = ord2f(((volatile unsigned *)scratch)[1]);
            if (((volatile unsigned *)scratch)[2]) a = z = __int_as_float(0x7fc00000);
            float scale, zp;
            quant_params<DT>(a, z, bits, scale, zp);
            qparams[0] = scale; qparams[1] = zp;
            scratch[0] = 0xffffffffu; scratch[1] = 0u; scratch[2] = 0u; scratch[3] = 0u;
        }
    }
}
__global__ void minmax_init_kernel(unsigned *scratch)
{
    scratch[0] = 0xffffffffu; scratch[1] = 0u; scratch[2] = 0u; scratch[3] = 0u;
}

template <int DT>
__global__ void __launch_bounds__(kQT) quantize_with_params_kernel(const void *x, int64_t numel, int bits,
                                                                   const float *qparams, void *y)
{
    const float scale = qparams[0], zp = qparams[1];
    const float qmax_r = qmax_rounded<DT>(bits);
    for (int64_t i = (int64_t)blockIdx.x * kQT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * kQT)
        store_elem<DT>(y, i, dequant_code<DT>(quant_code<DT>(load_elem<DT>(x, i), scale, zp, qmax_r), scale, zp));
}

// ---- packed rows (project-defined layout, DESIGN.md "packed rows") ---------------------
// One CTA per (row, batch).  bits in {1,2,4,8}: 8/bits codes per byte, first code in
// the least significant bits; bits in 9..16: one little-endian uint16 per code.
template <int DT>
__global__ void __launch_bounds__(kQT) quant_pack_kernel(QGParams p, int64_t row_bytes, uint8_t *codes, float *qparams)
{
    __shared__ float red[16];
    const int r = blockIdx.x, b = blockIdx.y;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    uint8_t *dst = codes + ((size_t)b * p.out_rows + r) * (size_t)row_bytes;
    float *qp = qparams + ((size_t)b * p.out_rows + r) * 2;
    if (r >= cnt) {
        if (p.zero_pad) {
            for (int64_t c = threadIdx.x; c < row_bytes; c += kQT) dst[c] = 0;
            if (threadIdx.x == 0) { qp[0] = 0.0f; qp[1] = 0.0f; }
        }
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const int64_t src = ((int64_t)b * p.S + tok) * p.n;
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        float v = load_elem<DT>(p.x[0], src + c);
        nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    block_minmax(mn, mx, nan, red);
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    float scale, zp;
    quant_params<DT>(mn, mx, bits, scale, zp);
    const float qmax_r = qmax_rounded<DT>(bits);
    if (threadIdx.x == 0) { qp[0] = scale; qp[1] = zp; }
    if (bits > 8) {
        for (int c = threadIdx.x; c < p.n; c += kQT) {
            float q = quant_code<DT>(load_elem<DT>(p.x[0], src + c), scale, zp, qmax_r);
            ((uint16_t *)dst)[c] = (uint16_t)(q == q ? (unsigned)q : 0u);
        }
        for (int64_t c = 2 * (int64_t)p.n + threadIdx.x; c < row_bytes; c += kQT) dst[c] = 0;
    } else {
        const int per = 8 / bits;
        const int nbytes = (p.n + per - 1) / per;
        for (int by = threadIdx.x; by < nbytes; by += kQT) {
            unsigned acc = 0;
            for (int e = 0; e < per; ++e) {
                const int c = by * per + e;
                if (c < p.n) {
                    float q = quant_code<DT>(load_elem<DT>(p.x[0], src + c), scale, zp, qmax_r);
                    acc |= (q == q ? (unsigned)q : 0u) << (e * bits);
                }
            }
            dst[by] = (uint8_t)acc;
        }
        for (int64_t c = nbytes + threadIdx.x; c < row_bytes; c += kQT) dst[c] = 0;
    }
}

// Vector path of quant_pack_kernel: the K3 structure (row in registers, 128-bit streaming loads, one block
// min/max) with a packing epilogue.  A lane's 16-byte vector becomes EPV codes = EPV*bits/8 bytes, written as
// one 1/2/4/8-byte store; consecutive lanes write consecutive bytes, so a warp's store is one contiguous
// 32..256-byte run.  Needs EPV*bits to be a whole number of bytes (everything but fp32 with 1-bit codes).
template <int DT, int NV>
__global__ void __launch_bounds__(kQT) quant_pack_vec_kernel(QGParams p, int64_t row_bytes, uint8_t *codes, float *qparams)
{
    constexpr int EPV = Vec16<DT>::kElems;
    __shared__ float red[16];
    const int r = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    pdl_trigger();
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int nvec = p.n / EPV;
    uint8_t *dst = codes + ((size_t)b * p.out_rows + r) * (size_t)row_bytes;
    float *qp = qparams + ((size_t)b * p.out_rows + r) * 2;
    if (r >= cnt) {
        for (int64_t c = threadIdx.x; c < row_bytes / 16; c += kQT) ((uint4 *)dst)[c] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { qp[0] = 0.0f; qp[1] = 0.0f; }
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
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (threadIdx.x + k * kQT < nvec) {
            float f[EPV];
            Vec16<DT>::unpack(raw[k], f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) { nan |= (f[e] != f[e]); mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
        }
    }
    block_minmax(mn, mx, nan, red);
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    RowQuant<DT> rq;
    rq.init(mn, mx, bits);
    if (threadIdx.x == 0) { qp[0] = rq.scale; qp[1] = rq.zp; }
    const int vbytes = bits > 8 ? 2 * EPV : EPV * bits / 8;       // bytes of codes per 16-byte input vector
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = threadIdx.x + k * kQT;
        if (v < nvec) {
            float f[EPV], cf[EPV];
            Vec16<DT>::unpack(raw[k], f);
            if (rq.fast) rq.codes_fast(f, cf);
            else {
#pragma unroll
                for (int e = 0; e < EPV; ++e) cf[e] = rq.code_slow(f[e]);
            }
            unsigned long long lo = 0, hi = 0;                    // up to 16 bytes of codes
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float qf = cf[e];
                const unsigned long long c = qf == qf ? (unsigned long long)(unsigned)qf : 0ull;
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
    // zero the padding between the last code byte and row_bytes
    for (int64_t c = (int64_t)nvec * vbytes + threadIdx.x; c < row_bytes; c += kQT) dst[c] = 0;
}

template <int DT>
__global__ void __launch_bounds__(kQT) unpack_dequant_kernel(QGParams p, int64_t row_bytes, const uint8_t *codes,
                                                             const float *qparams)
{
    const int r = blockIdx.x, b = blockIdx.y;
    const int cnt = p.row_stats ? p.row_stats[b].n_selected : p.S;
    const int64_t dst = ((int64_t)b * p.out_rows + r) * p.n;
    if (r >= cnt) {
        if (p.zero_pad) for (int c = threadIdx.x; c < p.n; c += kQT) store_elem<DT>(p.y[0], dst + c, 0.0f);
        return;
    }
    const int tok = p.sel_idx ? p.sel_idx[(int64_t)b * p.S + r] : r;
    const uint8_t *src = codes + ((size_t)b * p.out_rows + r) * (size_t)row_bytes;
    const float scale = qparams[((size_t)b * p.out_rows + r) * 2], zp = qparams[((size_t)b * p.out_rows + r) * 2 + 1];
    const int bits = pick_bits(p, p.labels[(int64_t)b * p.S + tok]);
    for (int c = threadIdx.x; c < p.n; c += kQT) {
        unsigned q;
        if (bits > 8) q = ((const uint16_t *)src)[c];
        else { const int per = 8 / bits; q = (src[c / per] >> ((c % per) * bits)) & ((1u << bits) - 1u); }
        store_elem<DT>(p.y[0], dst + c, dequant_code<DT>((float)q, scale, zp));
    }
}

} // namespace rtkv

using namespace rtkv;

extern "C" int rtkv_debug_bf16_packed(int on)
{
    const int v = on ? 1 : 0;
    RTKV_CUDA(cudaMemcpyToSymbol(g_bf16_packed, &v, sizeof(int)));
    return RTKV_OK;
}

extern "C" int rtkv_quant_gather(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                 const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                 const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                 void *Y0, void *Y1, void *stream)
{
    RTKV_CHECK_ARG(X0 && Y0 && (!X1 || Y1), "rtkv_quant_gather: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_quant_gather: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_gather: bad shape");
    RTKV_CHECK_ARG(B <= 65535, "rtkv_quant_gather: B > 65535");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_gather: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_quant_gather: out_rows > S without sel_idx");
    int rc = rtkv_check_bits(labels ? bits : nullptr, dtype, "rtkv_quant_gather");
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    RTKV_DISPATCH(dtype, return launch_qg<DT>(p, X1 ? 2 : 1, (cudaStream_t)stream, out_rows, B));
    return RTKV_OK;
}

extern "C" int rtkv_quant_gather_range(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                       const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                       const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                       void *Y0, void *Y1, int b_begin, int b_count, int row_begin, int row_count,
                                       void *stream)
{
    RTKV_CHECK_ARG(X0 && Y0 && (!X1 || Y1), "rtkv_quant_gather_range: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_quant_gather_range: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_gather_range: bad shape");
    RTKV_CHECK_ARG(b_begin >= 0 && b_count >= 0 && b_begin + b_count <= B && b_count <= 65535 && row_begin >= 0 &&
                   row_count >= 0 && row_begin + row_count <= out_rows, "rtkv_quant_gather_range: bad range");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_gather_range: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_quant_gather_range: out_rows > S without sel_idx");
    int rc = rtkv_check_bits(labels ? bits : nullptr, dtype, "rtkv_quant_gather_range");
    if (rc) return rc;
    if (row_count == 0 || b_count == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    p.row_begin = row_begin; p.b_begin = b_begin;
    RTKV_DISPATCH(dtype, return launch_qg<DT>(p, X1 ? 2 : 1, (cudaStream_t)stream, row_count, b_count));
    return RTKV_OK;
}

extern "C" int rtkv_row_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                               const int32_t *sel_idx, const rtkv_row_stats *row_stats, int out_rows,
                               int negate_max, float *minmax, void *stream)
{
    RTKV_CHECK_ARG(X0 && minmax, "rtkv_row_minmax: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype) && B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "rtkv_row_minmax: bad shape");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_row_minmax: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_row_minmax: out_rows > S without sel_idx");
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, nullptr, nullptr, out_rows, 0, nullptr, nullptr);
    p.neg_max = negate_max != 0;
    dim3 grid(out_rows, B, X1 ? 2 : 1);
    const int epv = 16 / dtype_bytes(dtype);
    const bool vec = n % epv == 0 && (uintptr_t)X0 % 16 == 0 && (!X1 || (uintptr_t)X1 % 16 == 0);
    if (vec) { RTKV_DISPATCH(dtype, (row_minmax_vec_kernel<DT><<<grid, kQT, 0, (cudaStream_t)stream>>>(p, minmax))); }
    else { RTKV_DISPATCH(dtype, (row_minmax_kernel<DT><<<grid, kQT, 0, (cudaStream_t)stream>>>(p, minmax))); }
    RTKV_LAUNCH_CHECK("row_minmax_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_quant_gather_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                        const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                        const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                        const float *minmax, int negated_max, void *Y0, void *Y1, void *stream)
{
    RTKV_CHECK_ARG(X0 && Y0 && (!X1 || Y1) && labels && bits && minmax, "rtkv_quant_gather_minmax: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_quant_gather_minmax: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_gather_minmax: bad shape");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_gather_minmax: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_quant_gather_minmax: out_rows > S without sel_idx");
    int rc = rtkv_check_bits(bits, dtype, "rtkv_quant_gather_minmax");
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    p.ext_minmax = minmax; p.neg_max = negated_max != 0;
    RTKV_DISPATCH(dtype, return launch_qg<DT>(p, X1 ? 2 : 1, (cudaStream_t)stream, out_rows, B));
    return RTKV_OK;
}

extern "C" int rtkv_quant_gather_token_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                              const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                              const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                              const float *minmax_by_token, void *Y0, void *Y1, void *stream)
{
    RTKV_CHECK_ARG(X0 && Y0 && (!X1 || Y1) && labels && bits && minmax_by_token, "rtkv_quant_gather_token_minmax: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_quant_gather_token_minmax: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_gather_token_minmax: bad shape");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_gather_token_minmax: sel_idx needs row_stats");
    RTKV_CHECK_ARG(sel_idx || out_rows <= S, "rtkv_quant_gather_token_minmax: out_rows > S without sel_idx");
    int rc = rtkv_check_bits(bits, dtype, "rtkv_quant_gather_token_minmax");
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    p.ext_minmax = minmax_by_token; p.ext_by_token = 1;
    const int epv = 16 / dtype_bytes(dtype);
    const int nt = X1 ? 2 : 1;
    const bool aligned = (uintptr_t)X0 % 16 == 0 && (uintptr_t)Y0 % 16 == 0 && (!X1 || ((uintptr_t)X1 % 16 == 0 && (uintptr_t)Y1 % 16 == 0));
    const int nvec = n % epv == 0 ? n / epv : 0;
    int shift = -1;
    for (int s = 3; s <= 7; ++s)                  // 8..128 threads per row slice, 4 vectors each
        if (nvec == (4 << s)) shift = s;
    if (aligned && shift >= 0) {
        const long long threads = ((long long)B * out_rows * nt) << shift;
        const long long ctas = (threads + kQT - 1) / kQT;
        RTKV_CHECK_ARG(ctas < (1ll << 31), "rtkv_quant_gather_token_minmax: too many rows");
        RTKV_DISPATCH(dtype, RTKV_CUDA(launch_kernel(quant_gather_flat_kernel<DT>, dim3((unsigned)ctas), dim3(kQT), 0,
                                                     (cudaStream_t)stream, 1, p, shift, nt)));
        return RTKV_OK;
    }
    RTKV_DISPATCH(dtype, return launch_qg<DT>(p, nt, (cudaStream_t)stream, out_rows, B));
    return RTKV_OK;
}

// smallest CTA whose threads hold the row slice in <= 8 vectors each
template <int DT> static void (*pick_p2p_kernel(int nvec, int *threads))(QGParams)
{
    if (nvec <= 64 * 4) { *threads = 64; return quant_gather_p2p_kernel<DT, 4, 64>; }
    if (nvec <= 64 * 8) { *threads = 64; return quant_gather_p2p_kernel<DT, 8, 64>; }
    if (nvec <= 128 * 8) { *threads = 128; return quant_gather_p2p_kernel<DT, 8, 128>; }
    *threads = 256;
    return quant_gather_p2p_kernel<DT, kMaxVec, 256>;
}

// L2 warm-up: one prefetch.global.L2 per 128-byte line of [p, p+bytes).  Meant to run on a side stream
// while the latency-bound selection kernel (K2, one SM) leaves HBM idle: the first rows the quantise
// kernel will read are then L2 hits.
__global__ void prefetch_l2_kernel(const char *p, size_t lines)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
}

extern "C" int rtkv_prefetch_l2(const void *p, size_t bytes, void *stream)
{
    RTKV_CHECK_ARG(p || bytes == 0, "rtkv_prefetch_l2: null pointer");
    const size_t lines = bytes / 128;
    if (lines == 0) return RTKV_OK;
    const size_t ctas = (lines + 255) / 256;
    prefetch_l2_kernel<<<(unsigned)(ctas < 148 * 16 ? ctas : 148 * 16), 256, 0, (cudaStream_t)stream>>>((const char *)p, lines);
    RTKV_LAUNCH_CHECK("prefetch_l2_kernel");
    return RTKV_OK;
}

extern "C" int64_t rtkv_p2p_mailbox_bytes(int B, int out_rows, int world)
{
    if (B <= 0 || out_rows <= 0 || world <= 0 || world > 8) return -1;
    return (int64_t)2 /*parity*/ * 2 /*K,V*/ * B * out_rows * world * 16;
}

extern "C" int rtkv_quant_gather_p2p(const void *X0, const void *X1, int dtype, int B, int S, int n,
                                     const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                     const uint8_t *labels, const int32_t *bits, int out_rows, int zero_pad,
                                     const uint64_t *mailbox_ptrs, int rank, int world, uint32_t *state,
                                     void *Y0, void *Y1, void *stream)
{
    RTKV_CHECK_ARG(X0 && X1 && Y0 && Y1 && sel_idx && row_stats && labels && bits && mailbox_ptrs && state,
                   "rtkv_quant_gather_p2p: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype) && B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows > 0, "rtkv_quant_gather_p2p: bad shape");
    RTKV_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "rtkv_quant_gather_p2p: bad rank/world");
    int rc = rtkv_check_bits(bits, dtype, "rtkv_quant_gather_p2p");
    if (rc) return rc;
    const int epv = 16 / dtype_bytes(dtype);
    if (n % epv != 0 || n > kMaxVec * kQT * epv || (uintptr_t)X0 % 16 || (uintptr_t)X1 % 16 || (uintptr_t)Y0 % 16 ||
        (uintptr_t)Y1 % 16) {
        set_error("rtkv_quant_gather_p2p: rows must be 16-byte aligned vectors of at most %d elements", kMaxVec * kQT * epv);
        return RTKV_EUNSUPPORTED;
    }
    QGParams p;
    fill_params(p, X0, X1, B, S, n, sel_idx, row_stats, labels, bits, out_rows, zero_pad, Y0, Y1);
    for (int i = 0; i < world; ++i) p.mbox[i] = mailbox_ptrs[i];
    p.p2p_state = state; p.rank = rank; p.world = world; p.total_ctas = out_rows * B * 2;
    dim3 grid(out_rows, B, 2);
    cudaStream_t st = (cudaStream_t)stream;
    void (*kern)(QGParams) = nullptr;
    int threads = 256;
    RTKV_DISPATCH(dtype, kern = pick_p2p_kernel<DT>(n / epv, &threads));
    RTKV_CUDA(launch_kernel(kern, grid, dim3(threads), 0, st, 1, p));
    return RTKV_OK;
}

extern "C" int rtkv_quant_params(const void *x, int dtype, int64_t numel, int bits, float *qparams,
                                 void *scratch, void *stream)
{
    RTKV_CHECK_ARG(x && qparams && scratch && numel > 0 && dtype_ok(dtype), "rtkv_quant_params: bad argument");
    int32_t b3[3] = {bits, bits, bits};
    int rc = rtkv_check_bits(b3, dtype, "rtkv_quant_params");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *sc = (unsigned *)scratch + 16; // inside the 256-byte scratch header
    minmax_init_kernel<<<1, 1, 0, st>>>(sc);
    int64_t ctas = (numel + kQT * 8 - 1) / (kQT * 8);
    int grid = (int)(ctas < 148 * 8 ? ctas : 148 * 8);
    RTKV_DISPATCH(dtype, (tensor_minmax_kernel<DT><<<grid, kQT, 0, st>>>(x, numel, bits, sc, qparams)));
    RTKV_LAUNCH_CHECK("tensor_minmax_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_quantize_with_params(const void *x, int dtype, int64_t numel, int bits,
                                         const float *qparams, void *y, void *stream)
{
    RTKV_CHECK_ARG(x && y && qparams && numel > 0 && dtype_ok(dtype), "rtkv_quantize_with_params: bad argument");
    int32_t b3[3] = {bits, bits, bits};
    int rc = rtkv_check_bits(b3, dtype, "rtkv_quantize_with_params");
    if (rc) return rc;
    int64_t ctas = (numel + kQT * 4 - 1) / (kQT * 4);
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    RTKV_DISPATCH(dtype, (quantize_with_params_kernel<DT><<<grid, kQT, 0, (cudaStream_t)stream>>>(x, numel, bits, qparams, y)));
    RTKV_LAUNCH_CHECK("quantize_with_params_kernel");
    return RTKV_OK;
}

extern "C" int64_t rtkv_packed_row_bytes(int n, int bits)
{
    if (n <= 0 || bits < 1 || bits > 16) return -1;
    int64_t raw = bits > 8 ? 2 * (int64_t)n : ((int64_t)n + (8 / bits) - 1) / (8 / bits);
    return (raw + 15) / 16 * 16;
}

template <int DT> static void (*pick_pack_kernel(int nv))(QGParams, int64_t, uint8_t *, float *)
{
    if (nv <= 1) return quant_pack_vec_kernel<DT, 1>;
    if (nv <= 2) return quant_pack_vec_kernel<DT, 2>;
    if (nv <= 4) return quant_pack_vec_kernel<DT, 4>;
    if (nv <= 5) return quant_pack_vec_kernel<DT, 5>;
    return quant_pack_vec_kernel<DT, kMaxVec>;
}

static int check_pack_bits(const int32_t *bits, const char *who)
{
    for (int c = 0; c < 3; ++c) {
        int b = bits[c];
        if (!(b == 1 || b == 2 || b == 4 || b == 8 || (b > 8 && b <= 16))) {
            set_error("%s: packed rows need bits in {1,2,4,8,9..16}, got %d", who, b);
            return RTKV_EUNSUPPORTED;
        }
    }
    return RTKV_OK;
}

extern "C" int rtkv_quant_pack(const void *X, int dtype, int B, int S, int n,
                               const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                               const uint8_t *labels, const int32_t *bits, int out_rows, int64_t row_bytes,
                               uint8_t *codes, float *qparams, void *stream)
{
    RTKV_CHECK_ARG(X && labels && bits && codes && qparams, "rtkv_quant_pack: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype) && B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "rtkv_quant_pack: bad shape");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_quant_pack: sel_idx needs row_stats");
    int rc = rtkv_check_bits(bits, dtype, "rtkv_quant_pack");
    if (!rc) rc = check_pack_bits(bits, "rtkv_quant_pack");
    if (rc) return rc;
    for (int c = 0; c < 3; ++c)
        if (dtype == RTKV_BF16 && bits[c] == 16) {  // clamp bound rnd(65535) = 65536 has no 16-bit code
            set_error("rtkv_quant_pack: 16-bit codes are not representable for bf16 rows");
            return RTKV_EUNSUPPORTED;
        }
    int mb = bits[0] > bits[1] ? bits[0] : bits[1]; mb = mb > bits[2] ? mb : bits[2];
    RTKV_CHECK_ARG(row_bytes >= rtkv_packed_row_bytes(n, mb), "rtkv_quant_pack: row_bytes too small");
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, X, nullptr, B, S, n, sel_idx, row_stats, labels, bits, out_rows, 1, nullptr, nullptr);
    dim3 grid(out_rows, B);
    const int epv = 16 / dtype_bytes(dtype);
    int minb = bits[0] < bits[1] ? bits[0] : bits[1]; minb = minb < bits[2] ? minb : bits[2];
    const bool vec = n % epv == 0 && n <= kMaxVec * kQT * epv && (uintptr_t)X % 16 == 0 && (uintptr_t)codes % 16 == 0 &&
                     row_bytes % 16 == 0 && (epv * minb) % 8 == 0;
    if (vec) {
        const int nv = (n / epv + kQT - 1) / kQT;
        void (*kern)(QGParams, int64_t, uint8_t *, float *) = nullptr;
        RTKV_DISPATCH(dtype, kern = pick_pack_kernel<DT>(nv));
        RTKV_CUDA(launch_kernel(kern, grid, dim3(kQT), 0, (cudaStream_t)stream, 1, p, row_bytes, codes, qparams));
    } else {
        RTKV_DISPATCH(dtype, (quant_pack_kernel<DT><<<grid, kQT, 0, (cudaStream_t)stream>>>(p, row_bytes, codes, qparams)));
    }
    RTKV_LAUNCH_CHECK("quant_pack_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_unpack_dequant(const uint8_t *codes, const float *qparams, int dtype, int B, int S, int n,
                                   const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                                   const uint8_t *labels, const int32_t *bits, int out_rows, int64_t row_bytes,
                                   void *Y, void *stream)
{
    RTKV_CHECK_ARG(codes && qparams && labels && bits && Y, "rtkv_unpack_dequant: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype) && B > 0 && B <= 65535 && S > 0 && n > 0 && out_rows >= 0, "rtkv_unpack_dequant: bad shape");
    RTKV_CHECK_ARG(!sel_idx || row_stats, "rtkv_unpack_dequant: sel_idx needs row_stats");
    int rc = check_pack_bits(bits, "rtkv_unpack_dequant");
    if (rc) return rc;
    if (out_rows == 0) return RTKV_OK;
    QGParams p;
    fill_params(p, nullptr, nullptr, B, S, n, sel_idx, row_stats, labels, bits, out_rows, 1, Y, nullptr);
    dim3 grid(out_rows, B);
    RTKV_DISPATCH(dtype, (unpack_dequant_kernel<DT><<<grid, kQT, 0, (cudaStream_t)stream>>>(p, row_bytes, codes, qparams)));
    RTKV_LAUNCH_CHECK("unpack_dequant_kernel");
    return RTKV_OK;
}
