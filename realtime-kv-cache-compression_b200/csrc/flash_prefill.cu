// flash_prefill.cu — prefill attention on the 5th-generation tensor cores, written for the compression path's two
// neighbours:
//   upstream   src/models/modified_llama.py:88-94    softmax(Q K^T / sqrt(d) + causal mask), materialised [B,H,S,S]
//              -> here: no S x S tensor; the kernel returns the attention output, the row log-sum-exp AND the
//              prompt-column mass  head_sum[b,h,i] = sum_{j<P, j<=i} softmax_ij  that token_importance.py:41-45 takes
//              from the materialised probabilities (it is a by-product of the first key tile: free).
//   downstream src/models/modified_llama.py:125-142  softmax(Q K_sel^T / sqrt(d) + mask[..., :S']) V_sel over the
//              S' kept (fake-quantised) tokens, with the reference's truncated mask: query i sees kept slots j <= i
//              -> the same kernel with kv_len = S' < q_len.
//
// Five kernels sit behind rtkv_flash_prefill (rtkv_debug_flash_variant; DESIGN.md §5 K-flash has the measurements that led from
// one to the next).  By default the entry launches flash_prefill_kernel (described next) below 3072 keys and
// flash_prefill_mc_kernel (alternating key tiles, one accumulator, P in tensor memory, three K/V stages, separate issuing
// threads for the two products; described at its definition) from there on.
//
// flash_prefill_kernel: one CTA = 128 query rows of one (batch row, head); keys/values stream through in tiles of 128:
//   warp 0      TMA producer: Q tile once, K and V tiles through 2-stage rings (4-D tensor maps: any [B,H,S,128]
//               view with unit stride in the last dim, e.g. the transposed view of a [B,S,H,128] projection)
//   warp 1      MMA issuer (one thread): S_t = Q K_t^T (tcgen05.mma 128x128x16, both operands K-major, fp32 in TMEM,
//               two S buffers so that S_{t+1} is computed while softmax works on S_t), then O += P_t V_t with P from
//               shared memory (K-major, written by the softmax warps in the 128-byte swizzle) and V as an MN-major
//               operand straight from the TMA tile (no transpose anywhere)
//   warps 2..9  softmax, two warpgroups: warpgroup g owns key columns [64g, 64g+64) of the tile for all 128 query rows
//               (a thread = one row x 64 columns; the two halves of a row exchange only their partial row max, through
//               shared memory, once per tile; partial row sums are added at the very end).  tcgen05.ld of the logits,
//               running max with LAZY rescaling (the accumulator in TMEM is only rescaled when the row max grows by more
//               than 2^8, which keeps P <= 256 and makes the TMEM read-modify-write of O rare), exp2, bf16 P half-tile to
//               shared memory (double buffered, so softmax of tile t+1 runs while the tensor pipe does P_t V_t)
// Numerics: logits and row sums in fp32, P rounded to the input dtype for the second MMA (flash-attention practice);
// O accumulates in fp32 in TMEM; LSE = ln2 * (m + log2(l)).
#include "umma.cuh"

namespace rtkv {

constexpr int kFT = 128;                         // tile: 128 query rows x 128 keys, head_dim 128
constexpr int kFTileBytes = kFT * 128 * 2;       // 32 KiB: one Q / K / V / P tile (bf16 or f16)
constexpr int kFHalf = kFTileBytes / 2;          // the two 64-column (128-byte) halves of a tile
constexpr int kFThreads = 64 + 256;              // TMA warp, MMA warp, 8 softmax warps (two warpgroups)
// Q, K[2], V[2], P[2] + barriers + row exchange + alignment slack: 232448 bytes = the 227 KB a CTA can have
constexpr size_t kFSmem = 768 + 7 * (size_t)kFTileBytes + 256 + 4 * kFT * sizeof(float);
constexpr size_t kIHSmem = 768 + 7 * (size_t)kFTileBytes + 256;   // the epilogue exchange reuses the P buffers

struct FlashParams {
    void *out;                 // [B,H,Sq,128] elements of the input dtype, strides below (in elements)
    int64_t o_sb, o_sh, o_ss;
    float *lse;                // [B,H,Sq] or NULL
    float *head_sums;          // [B,H,Sq] or NULL: prompt-column mass
    int B, H, Hkv, Sq, Skv, P, causal, is_bf16;
    float scale_log2;          // sm_scale * log2(e)
    uint32_t idesc_qk, idesc_pv;
    int skew;                  // alternating-tile kernel: cycles the odd warpgroup starts late (0: none)
    long long *prof;           // debug: per-phase cycle counters of CTA 0 (rtkv_debug_flash_profile), else NULL
};

template <bool BF16> __device__ __forceinline__ uint32_t pack2(float a, float b)
{
    if constexpr (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// 2^x on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, degree-4 polynomial for 2^f
// (max relative error 4.2e-5, far below the bf16/f16 rounding of P and the 2e-3 tolerance of LSE / prompt mass), 2^n by adding
// n to the exponent field.  The exponentials are as expensive as the tensor work on this machine (16 MUFU results per clock
// and SM against 8192 tensor FLOP: 1024 cycles each per 128 x 128 tile), so a share of them runs here while the MUFU pipe
// is busy with the rest (the FlashAttention-4 recipe).
__device__ __forceinline__ float exp2_fma(float x)
{
    x = fmaxf(x, -126.0f);
    const float xr = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits of xr hold round(x)
    const float f = x - (xr - 12582912.0f);
    float p = __fmaf_rn(f, 0.009618129f, 0.055504109f);
    p = __fmaf_rn(p, f, 0.240226507f);
    p = __fmaf_rn(p, f, 0.693147181f);
    p = __fmaf_rn(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

// One thread's 64 logits of a tile -> p = 2^(s*scale - m), row sum, P half-row (input dtype, K-major, 128-byte swizzle) in
// shared memory.  SLOW: the tile needs masking (diagonal / key tail) or is the first tile (prompt-column sum); every
// other tile runs the branch-free body: FFMA + MUFU.EX2 + FADD per element and one pack per pair.
template <bool BF16, bool SLOW, int PE>
__device__ __forceinline__ void exp_store(const uint32_t (&raw)[2][32], float scale_log2, float m_used, uint32_t prow, int row,
                                          int j_base, int i, int Skv, int causal, int P, bool first, float &rs, float &ps)
{
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
            float p0, p1;
            if (!SLOW && PE > 0 && ((e >> 1) % (PE > 0 ? PE : 1)) == PE - 1) {     // this pair on the FMA pipe
                p0 = exp2_fma(__fmaf_rn(__uint_as_float(raw[c][e]), scale_log2, -m_used));
                p1 = exp2_fma(__fmaf_rn(__uint_as_float(raw[c][e + 1]), scale_log2, -m_used));
            } else {
                p0 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e]), scale_log2, -m_used));
                p1 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e + 1]), scale_log2, -m_used));
            }
            if constexpr (SLOW) {
                const int j = j_base + 32 * c + e;
                if (j >= Skv || (causal && j > i)) p0 = 0.0f;
                if (j + 1 >= Skv || (causal && j + 1 > i)) p1 = 0.0f;
                if (first) ps += (j < P ? p0 : 0.0f) + (j + 1 < P ? p1 : 0.0f);
            }
            rs += p0 + p1;
            pk[e >> 1] = pack2<BF16>(p0, p1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)                              // four 16-byte chunks, XOR-swizzled by the row
            sts_v4(prow + (uint32_t)((((c * 4 + q) ^ (row & 7))) * 16), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
}

template <bool BF16, int PE, bool PROF = false>
__global__ void __launch_bounds__(kFThreads, 1)
flash_prefill_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, FlashParams p)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    // 128-byte swizzle needs 1024-byte aligned tiles; the declared alignment makes the round-up a no-op, the 768 spare bytes
    // cover a base that is only 256-byte aligned
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_q = smem, *smem_k = smem + kFTileBytes, *smem_v = smem + 3 * kFTileBytes, *smem_p = smem + 5 * kFTileBytes;
    uint64_t *bars = (uint64_t *)(smem + 7 * (size_t)kFTileBytes);
    uint64_t *qfull = bars, *kfull = bars + 1, *kempty = bars + 3, *vfull = bars + 5, *vempty = bars + 7;
    uint64_t *sfull = bars + 9, *sfree = bars + 11, *pfull = bars + 13, *pvdone = bars + 15;
    uint32_t *tmem_slot = (uint32_t *)(bars + 17);
    float *xch = (float *)((unsigned char *)bars + 256);          // [2 parities][2 warpgroups][128 rows] partial row max / sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (p.Sq + kFT - 1) / kFT;
    const int qt = n_qt - 1 - (int)(blockIdx.x % n_qt);          // heaviest (last) query tiles first
    const int h = (int)((blockIdx.x / n_qt) % p.H), b = (int)(blockIdx.x / ((size_t)n_qt * p.H));
    const int hk = h / (p.H / p.Hkv);
    int T = (p.Skv + kFT - 1) / kFT;                             // key tiles this query tile sees
    if (p.causal && qt + 1 < T) T = qt + 1;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(qfull), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1);
            mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1);
            mbar_init(smem_u32(&sfull[s]), 1); mbar_init(smem_u32(&sfree[s]), 8);
            mbar_init(smem_u32(&pfull[s]), 8); mbar_init(smem_u32(&pvdone[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // 512 TMEM columns: S0 [0,128), S1 [128,256), O [256,384)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t qf = smem_u32(qfull);
            mbar_expect_tx(qf, kFTileBytes);
            tma_load_4d(smem_u32(smem_q), &map_q, 0, qt * kFT, h, b, qf);
            tma_load_4d(smem_u32(smem_q) + kFHalf, &map_q, 64, qt * kFT, h, b, qf);
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                const uint32_t par = (uint32_t)(((t >> 1) & 1) ^ 1);
                mbar_wait(smem_u32(&kempty[s]), par);
                const uint32_t kf = smem_u32(&kfull[s]);
                mbar_expect_tx(kf, kFTileBytes);
                const uint32_t kd = smem_u32(smem_k + (size_t)s * kFTileBytes);
                tma_load_4d(kd, &map_k, 0, t * kFT, hk, b, kf);
                tma_load_4d(kd + kFHalf, &map_k, 64, t * kFT, hk, b, kf);
                mbar_wait(smem_u32(&vempty[s]), par);
                const uint32_t vf = smem_u32(&vfull[s]);
                mbar_expect_tx(vf, kFTileBytes);
                const uint32_t vd = smem_u32(smem_v + (size_t)s * kFTileBytes);
                tma_load_4d(vd, &map_v, 0, t * kFT, hk, b, vf);
                tma_load_4d(vd + kFHalf, &map_v, 64, t * kFT, hk, b, vf);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // descriptors are built once and advanced by one add per MMA: issuing, not executing, was the bottleneck (umma.cuh)
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem_q)), kd0 = umma_desc_sw128(smem_u32(smem_k));
            const uint64_t pd0 = umma_desc_sw128(smem_u32(smem_p)), vd0 = umma_desc_sw128_mn(smem_u32(smem_v), kFHalf);
            const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
            auto issue_qk = [&](int t) {
                const int s = t & 1;
                mbar_wait(smem_u32(&kfull[s]), (uint32_t)((t >> 1) & 1));
                mbar_wait(smem_u32(&sfree[s]), (uint32_t)(((t >> 1) & 1) ^ 1));      // softmax has S_{t-2} in registers
                tc_fence_after();
                const uint64_t kd = umma_desc_advance(kd0, (uint32_t)(s * kFTileBytes));
                const uint32_t dcol = tmem_base + (uint32_t)(s * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = (uint32_t)(kq * kFHalf + kk * 32);       // 16 head-dim elements = 32 B
                        umma_f16(dcol, umma_desc_advance(qd0, off), umma_desc_advance(kd, off), idesc_qk, (uint32_t)((kq | kk) != 0));
                    }
                umma_commit(smem_u32(&kempty[s]));
                umma_commit(smem_u32(&sfull[s]));
            };
            mbar_wait(smem_u32(qfull), 0);
            issue_qk(0);
            long long mma_wait_p = 0, mma_t0 = 0;
            if constexpr (PROF) mma_t0 = clock64();
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) issue_qk(t + 1);                                      // overlaps softmax of tile t
                const int s = t & 1;
                long long w0 = 0;
                if constexpr (PROF) w0 = clock64();
                mbar_wait(smem_u32(&pfull[s]), (uint32_t)((t >> 1) & 1));            // P_t in shared memory, O rescaled
                if constexpr (PROF) mma_wait_p += clock64() - w0;
                mbar_wait(smem_u32(&vfull[s]), (uint32_t)((t >> 1) & 1));
                tc_fence_after();
                const uint64_t vd = umma_desc_advance(vd0, (uint32_t)(s * kFTileBytes));
                const uint64_t pd = umma_desc_advance(pd0, (uint32_t)(s * kFTileBytes));
                const uint32_t ocol = tmem_base + 2 * kFT;
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)          // keys 0..63 / 64..127: the two halves of the P tile
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {    // 16 keys per MMA
                        const uint32_t poff = (uint32_t)(kq * kFHalf + kk * 32);
                        const uint32_t voff = (uint32_t)((kq * 64 + kk * 16) * 128);  // 16 key rows of 128 B in each head-dim half
                        umma_f16(ocol, umma_desc_advance(pd, poff), umma_desc_advance(vd, voff), idesc_pv,
                                 (uint32_t)((t | kq | kk) != 0));
                    }
                umma_commit(smem_u32(&vempty[s]));
                umma_commit(smem_u32(&pvdone[s]));
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0 && p.prof) { p.prof[32] = mma_wait_p; p.prof[33] = clock64() - mma_t0; p.prof[34] = 0; p.prof[35] = T; }
            }
        }
    } else {
        // ===== softmax + epilogue: thread = one query row x 64 key columns (warpgroup g = column half g) =====
        const int quad = warp & 3;
        const int g = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const int i = qt * kFT + row;                          // query index
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        float m_used = -INFINITY, l = 0.0f, mass_acc = 0.0f;   // l, mass_acc: this thread's column half only
        const uint32_t xch_u32 = smem_u32(xch);
        // PROF (tools/profile_flash_phases.py): cycles per phase of the tile loop, summed over the tiles of CTA 0 by lane 0 of the
        // first warp of each warpgroup
        long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_prev = 0;
        auto tick = [&](int k) {
            if constexpr (PROF) {
                const long long now = clock64();
                prof_acc[k] += now - prof_prev;
                prof_prev = now;
            }
        };
        if constexpr (PROF) prof_prev = clock64();
        for (int t = 0; t < T; ++t) {
            const int s = t & 1;
            tick(0);
            mbar_wait(smem_u32(&sfull[s]), (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            tick(1);
            uint32_t raw[2][32];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld32_issue(lane_base + (uint32_t)(s * kFT + 64 * g + 32 * c), raw[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sfree[s]));
            tick(2);
            // row max of this thread's 64 logits (raw units; masked entries only where a tile needs masking)
            const int j0 = t * kFT + 64 * g;
            const bool need_mask = (p.causal && t * kFT + kFT - 1 > qt * kFT) || (t * kFT + kFT > p.Skv);
            float mx = -INFINITY;
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = j0 + 32 * c + e;
                        const bool dead = j >= p.Skv || (p.causal && j > i);
                        mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(raw[c][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(raw[c][e]));
            }
            // the other half of the row: one float through shared memory, one named barrier per tile
            const uint32_t xm = xch_u32 + (uint32_t)((t & 1) * 2 * kFT * 4);
            if constexpr (PROF) { if (mx == 12345.678f) prof_acc[7] += 1; }   // force the max before the clock read
            tick(3);
            sts_f32(xm + (uint32_t)((g * kFT + row) * 4), mx);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(mx, lds_f32(xm + (uint32_t)(((g ^ 1) * kFT + row) * 4))) * p.scale_log2;   // log2 domain from here
            // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
            float alpha = 1.0f;
            bool rescale = false;
            if (mx > m_used + 8.0f) {
                rescale = m_used != -INFINITY;
                alpha = rescale ? ex2_approx(m_used - mx) : 0.0f;
                m_used = mx;
            }
            if (__any_sync(0xffffffffu, rescale)) {              // warp-uniform: tcgen05.ld/st are warp-wide; rare
                mbar_wait(smem_u32(&pvdone[(t - 1) & 1]), (uint32_t)(((t - 1) >> 1) & 1));   // P_{t-1} V_{t-1} has retired: O stable
                tc_fence_after();
                const float a = rescale ? alpha : 1.0f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t o[32];
                    tmem_ld32(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a);
                    tmem_st32(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o);
                }
                tmem_st_wait();
            }
            if (rescale) { l *= alpha; mass_acc *= alpha; }
            tick(4);
            // P buffer (t & 1) is free once P_{t-2} V_{t-2} has retired
            mbar_wait(smem_u32(&pvdone[s]), (uint32_t)(((t >> 1) & 1) ^ 1));
            tick(5);
            // p = 2^(s*scale - m), row sum, prompt-column share, P half-tile to shared memory
            float rs = 0.0f, ps = 0.0f;
            const uint32_t prow = smem_u32(smem_p) + (uint32_t)(s * kFTileBytes + g * kFHalf + row * 128);
            if (need_mask || t == 0)
                exp_store<BF16, true, PE>(raw, p.scale_log2, m_used, prow, row, j0, i, p.Skv, p.causal, p.P, t == 0, rs, ps);
            else
                exp_store<BF16, false, PE>(raw, p.scale_log2, m_used, prow, row, j0, i, p.Skv, p.causal, p.P, false, rs, ps);
            l += rs;
            if (t == 0) mass_acc = ps;
            if constexpr (PROF) { if (l == 12345.678f) prof_acc[7] += 1; }
            tick(6);
            fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&pfull[s]));
            tick(7);
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && lane == 0 && quad == 2 && p.prof) {   // warps 2 and 6: the first warp of each warpgroup
                for (int k = 0; k < 8; ++k) p.prof[g * 16 + k] = prof_acc[k];
                p.prof[g * 16 + 8] = T;
            }
        }
        // ---- epilogue: total row sum over the two halves, O / l, LSE, prompt mass ----
        float *xl = xch + (size_t)(T & 1) * 2 * kFT;             // the slot the last tile did NOT use
        xl[g * kFT + row] = l;
        float *xp = xch + (size_t)((T & 1) ^ 1) * 2 * kFT;
        asm volatile("bar.sync 1, 256;" ::: "memory");           // everybody is past the last tile's max exchange
        xp[g * kFT + row] = mass_acc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += xl[(g ^ 1) * kFT + row];
        mass_acc += xp[(g ^ 1) * kFT + row];
        mbar_wait(smem_u32(&pvdone[(T - 1) & 1]), (uint32_t)(((T - 1) >> 1) & 1));
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool live = i < p.Sq;
        unsigned char *orow = (unsigned char *)p.out + ((size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)i * p.o_ss) * 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o);
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pack2<BF16>(__uint_as_float(o[8 * q + 2 * e]) * inv_l, __uint_as_float(o[8 * q + 2 * e + 1]) * inv_l);
                    *reinterpret_cast<uint4 *>(orow + (size_t)(64 * g + 32 * c + 8 * q) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        if (live && g == 0) {
            const size_t at = ((size_t)b * p.H + h) * p.Sq + i;
            if (p.lse) p.lse[at] = (m_used + log2f(l)) * 0.6931471805599453f;
            if (p.head_sums) p.head_sums[at] = mass_acc * inv_l;
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- independent key halves (rtkv_debug_flash_variant(1); measured slightly SLOWER: 855 vs 898 TFLOP/s at 32k) -----------
// Same CTA layout as flash_prefill_kernel, but the two softmax warpgroups no longer meet once per tile to agree on the row
// max (that barrier kept all eight warps in lockstep: the MUFU pipe idled while everybody loaded, maxed and stored).  Each
// warpgroup runs its OWN online softmax over its half of the keys of every tile (own running max, own row sum, own
// accumulator O_g in TMEM, own P half-tile and barriers), the second product is issued per half (4 MMAs of 16 keys into O_g),
// and the two partial results are merged once, in the epilogue: O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_g = 2^(m_g - M).
// Outcome: the per-tile barrier was not the limiter.  Per tile and scheduler the exponentials alone need 1024 MUFU cycles (64 per
// warp x 8 cycles x 2 warps) of the ~2470 a tile takes; the other phases (TMEM load, max, pack, stores) do not overlap them
// because both warps of a scheduler start every tile together.  Kept as a tested variant.
template <bool BF16>
__global__ void __launch_bounds__(kFThreads, 1)
flash_prefill_ih_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, FlashParams p)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    // 128-byte swizzle needs 1024-byte aligned tiles; the declared alignment makes the round-up a no-op, the 768 spare bytes
    // cover a base that is only 256-byte aligned
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_q = smem, *smem_k = smem + kFTileBytes, *smem_v = smem + 3 * kFTileBytes, *smem_p = smem + 5 * kFTileBytes;
    uint64_t *bars = (uint64_t *)(smem + 7 * (size_t)kFTileBytes);
    uint64_t *qfull = bars, *kfull = bars + 1, *kempty = bars + 3, *vfull = bars + 5, *vempty = bars + 7;
    uint64_t *sfull = bars + 9, *sfree = bars + 11, *pfull = bars + 13, *pvdone = bars + 17;   // pfull / pvdone: [half g][buffer s]
    uint32_t *tmem_slot = (uint32_t *)(bars + 21);
    float *xch = (float *)((unsigned char *)bars + 256);          // epilogue exchange: [3 values][2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (p.Sq + kFT - 1) / kFT;
    const int qt = n_qt - 1 - (int)(blockIdx.x % n_qt);          // heaviest (last) query tiles first
    const int h = (int)((blockIdx.x / n_qt) % p.H), b = (int)(blockIdx.x / ((size_t)n_qt * p.H));
    const int hk = h / (p.H / p.Hkv);
    int T = (p.Skv + kFT - 1) / kFT;                             // key tiles this query tile sees
    if (p.causal && qt + 1 < T) T = qt + 1;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(qfull), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1);
            mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1);
            mbar_init(smem_u32(&sfull[s]), 1); mbar_init(smem_u32(&sfree[s]), 8);
            for (int g = 0; g < 2; ++g) { mbar_init(smem_u32(&pfull[2 * g + s]), 4); mbar_init(smem_u32(&pvdone[2 * g + s]), 1); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // 512 TMEM columns: S0 [0,128), S1 [128,256), O_0 [256,384), O_1 [384,512) (one accumulator per key half)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t qf = smem_u32(qfull);
            mbar_expect_tx(qf, kFTileBytes);
            tma_load_4d(smem_u32(smem_q), &map_q, 0, qt * kFT, h, b, qf);
            tma_load_4d(smem_u32(smem_q) + kFHalf, &map_q, 64, qt * kFT, h, b, qf);
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                const uint32_t par = (uint32_t)(((t >> 1) & 1) ^ 1);
                mbar_wait(smem_u32(&kempty[s]), par);
                const uint32_t kf = smem_u32(&kfull[s]);
                mbar_expect_tx(kf, kFTileBytes);
                const uint32_t kd = smem_u32(smem_k + (size_t)s * kFTileBytes);
                tma_load_4d(kd, &map_k, 0, t * kFT, hk, b, kf);
                tma_load_4d(kd + kFHalf, &map_k, 64, t * kFT, hk, b, kf);
                mbar_wait(smem_u32(&vempty[s]), par);
                const uint32_t vf = smem_u32(&vfull[s]);
                mbar_expect_tx(vf, kFTileBytes);
                const uint32_t vd = smem_u32(smem_v + (size_t)s * kFTileBytes);
                tma_load_4d(vd, &map_v, 0, t * kFT, hk, b, vf);
                tma_load_4d(vd + kFHalf, &map_v, 64, t * kFT, hk, b, vf);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t qbase = smem_u32(smem_q);
            auto issue_qk = [&](int t) {
                const int s = t & 1;
                mbar_wait(smem_u32(&kfull[s]), (uint32_t)((t >> 1) & 1));
                mbar_wait(smem_u32(&sfree[s]), (uint32_t)(((t >> 1) & 1) ^ 1));      // softmax has S_{t-2} in registers
                tc_fence_after();
                const uint32_t kbase = smem_u32(smem_k + (size_t)s * kFTileBytes);
                const uint32_t dcol = tmem_base + (uint32_t)(s * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = (uint32_t)(kq * kFHalf + kk * 32);       // 16 head-dim elements = 32 B
                        umma_f16(dcol, umma_desc_sw128(qbase + off), umma_desc_sw128(kbase + off), p.idesc_qk, (uint32_t)((kq | kk) != 0));
                    }
                umma_commit(smem_u32(&kempty[s]));
                umma_commit(smem_u32(&sfull[s]));
            };
            mbar_wait(smem_u32(qfull), 0);
            issue_qk(0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) issue_qk(t + 1);                                      // overlaps softmax of tile t
                const int s = t & 1;
                mbar_wait(smem_u32(&vfull[s]), (uint32_t)((t >> 1) & 1));
                const uint32_t vbase = smem_u32(smem_v + (size_t)s * kFTileBytes);
                const uint32_t pbase = smem_u32(smem_p + (size_t)s * kFTileBytes);
#pragma unroll
                for (int g = 0; g < 2; ++g) {           // key half g (keys 64g .. 64g+63 of the tile) has its own accumulator O_g
                    mbar_wait(smem_u32(&pfull[2 * g + s]), (uint32_t)((t >> 1) & 1));   // P_t half g in shared memory, O_g rescaled
                    tc_fence_after();
                    const uint32_t ocol = tmem_base + (uint32_t)(2 * kFT + g * kFT);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {    // 16 keys per MMA
                        const uint32_t poff = (uint32_t)(g * kFHalf + kk * 32);
                        const uint32_t voff = (uint32_t)((g * 64 + kk * 16) * 128);
                        umma_f16(ocol, umma_desc_sw128(pbase + poff), umma_desc_sw128_mn(vbase + voff, kFHalf), p.idesc_pv,
                                 (uint32_t)((t | kk) != 0));
                    }
                    umma_commit(smem_u32(&pvdone[2 * g + s]));
                }
                umma_commit(smem_u32(&vempty[s]));
            }
        }
    } else {
        // ===== softmax + epilogue: thread = one query row x 64 key columns (warpgroup g = column half g) =====
        const int quad = warp & 3;
        const int g = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const int i = qt * kFT + row;                          // query index
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        float m_used = -INFINITY, l = 0.0f, mass_acc = 0.0f;   // l, mass_acc: this thread's column half only
        const uint32_t xch_u32 = smem_u32(xch);
        for (int t = 0; t < T; ++t) {
            const int s = t & 1;
            mbar_wait(smem_u32(&sfull[s]), (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            uint32_t raw[2][32];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld32_issue(lane_base + (uint32_t)(s * kFT + 64 * g + 32 * c), raw[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sfree[s]));
            // row max of this thread's 64 logits (raw units; masked entries only where a tile needs masking)
            const int j0 = t * kFT + 64 * g;
            const bool need_mask = (p.causal && t * kFT + kFT - 1 > qt * kFT) || (t * kFT + kFT > p.Skv);
            float mx = -INFINITY;
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = j0 + 32 * c + e;
                        const bool dead = j >= p.Skv || (p.causal && j > i);
                        mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(raw[c][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(raw[c][e]));
            }
            mx *= p.scale_log2;                                  // log2 domain from here; this half's own running max: no exchange
            // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
            float alpha = 1.0f;
            bool rescale = false;
            if (mx > m_used + 8.0f) {
                rescale = m_used != -INFINITY;
                alpha = rescale ? ex2_approx(m_used - mx) : 0.0f;
                m_used = mx;
            }
            if (__any_sync(0xffffffffu, rescale)) {              // warp-uniform: tcgen05.ld/st are warp-wide; rare
                mbar_wait(smem_u32(&pvdone[2 * g + ((t - 1) & 1)]), (uint32_t)(((t - 1) >> 1) & 1));   // P_{t-1} V_{t-1} (half g) retired
                tc_fence_after();
                const float a = rescale ? alpha : 1.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(lane_base + (uint32_t)(2 * kFT + g * kFT + 32 * c), o);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a);
                    tmem_st32(lane_base + (uint32_t)(2 * kFT + g * kFT + 32 * c), o);
                }
                tmem_st_wait();
            }
            if (rescale) { l *= alpha; mass_acc *= alpha; }
            // P buffer (t & 1) is free once P_{t-2} V_{t-2} has retired
            mbar_wait(smem_u32(&pvdone[2 * g + s]), (uint32_t)(((t >> 1) & 1) ^ 1));
            // p = 2^(s*scale - m), row sum, prompt-column share, P half-tile to shared memory
            float rs = 0.0f, ps = 0.0f;
            const uint32_t prow = smem_u32(smem_p) + (uint32_t)(s * kFTileBytes + g * kFHalf + row * 128);
            if (need_mask || t == 0)
                exp_store<BF16, true, 0>(raw, p.scale_log2, m_used, prow, row, j0, i, p.Skv, p.causal, p.P, t == 0, rs, ps);
            else
                exp_store<BF16, false, 0>(raw, p.scale_log2, m_used, prow, row, j0, i, p.Skv, p.causal, p.P, false, rs, ps);
            l += rs;
            if (t == 0) mass_acc = ps;
            fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&pfull[2 * g + s]));
        }
        // ---- epilogue: merge the two key halves' online softmaxes (each has its own max, sum and accumulator).  Both second
        // products of the last tile have retired first, so the P buffers are free: they carry the three floats per (row, half).
        mbar_wait(smem_u32(&pvdone[0 + ((T - 1) & 1)]), (uint32_t)(((T - 1) >> 1) & 1));
        mbar_wait(smem_u32(&pvdone[2 + ((T - 1) & 1)]), (uint32_t)(((T - 1) >> 1) & 1));
        tc_fence_after();
        float *xe = reinterpret_cast<float *>(smem_p);           // [3 values][2 halves][128 rows]
        xe[(0 * 2 + g) * kFT + row] = m_used;
        xe[(1 * 2 + g) * kFT + row] = l;
        xe[(2 * 2 + g) * kFT + row] = mass_acc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float m_o = xe[(0 * 2 + (g ^ 1)) * kFT + row], l_o = xe[(1 * 2 + (g ^ 1)) * kFT + row], mass_o = xe[(2 * 2 + (g ^ 1)) * kFT + row];
        const float M = fmaxf(m_used, m_o);
        const float a_me = ex2_approx(m_used - M), a_ot = ex2_approx(m_o - M);       // 2^(-inf) = 0 for a half that saw no key
        const float L = a_me * l + a_ot * l_o;
        const float inv_l = 1.0f / L;
        const float a0 = (g == 0 ? a_me : a_ot) * inv_l, a1 = (g == 0 ? a_ot : a_me) * inv_l;
        const bool live = i < p.Sq;
        unsigned char *orow = (unsigned char *)p.out + ((size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)i * p.o_ss) * 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                            // this thread writes output columns [64g, 64g + 64)
            uint32_t o0[32], o1[32];
            tmem_ld32_issue(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o0);
            tmem_ld32_issue(lane_base + (uint32_t)(3 * kFT + 64 * g + 32 * c), o1);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c0 = 8 * q + 2 * e;
                        w[e] = pack2<BF16>(__uint_as_float(o0[c0]) * a0 + __uint_as_float(o1[c0]) * a1,
                                           __uint_as_float(o0[c0 + 1]) * a0 + __uint_as_float(o1[c0 + 1]) * a1);
                    }
                    *reinterpret_cast<uint4 *>(orow + (size_t)(64 * g + 32 * c + 8 * q) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        if (live && g == 0) {
            const size_t at = ((size_t)b * p.H + h) * p.Sq + i;
            if (p.lse) p.lse[at] = (M + log2f(L)) * 0.6931471805599453f;
            if (p.head_sums) p.head_sums[at] = (a_me * mass_acc + a_ot * mass_o) * inv_l;
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- ping-pong variant (rtkv_debug_flash_variant(1); measured SLOWER than the kernel above: 577 vs 934 TFLOP/s at 32k) ----
// The single-tile kernel above leaves the tensor pipe idle while its softmax warps walk through their latency chain
// (TMEM load -> max -> exp -> P store): tensor pipe 38 % active.  The limiter turned out to be that latency chain per
// thread, not the hand-off to the tensor pipe: with one thread per row and all 128 columns the chain doubles and the
// ping-pong cannot hide it with one warp per (tile, scheduler).  Kept as the tested starting point for the FA4-style
// design (it proves the TMEM-resident P operand and the TS form of tcgen05.mma).  Here a CTA owns TWO query tiles (A, B: 256 rows) of one
// head; each has its own softmax warpgroup (a thread = one row x all 128 key columns: no exchange between threads at
// all), its own S and O accumulators in TMEM (4 x 128 columns = all 512), and the tensor pipe alternates between them:
// while warpgroup A runs softmax on S_A(t), the MMA thread issues P_B(t-1) V, Q_B K(t)^T, ...  P is written back to TMEM
// over the first 64 columns of its S (two bf16/f16 per 32-bit column) and used as the A operand of the second product
// straight from there (tcgen05.mma with A in tensor memory): no shared memory for P, which is what makes room for the
// second Q tile.  K/V tiles are shared by both query tiles (half the L2 -> shared-memory traffic per FLOP).
constexpr int kPPThreads = 64 + 256;
constexpr size_t kPPSmem = 768 + 6 * (size_t)kFTileBytes + 256;          // Q_A, Q_B, K[2], V[2] + barriers

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(kPPThreads, 1)
flash_prefill_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, FlashParams p)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_q = smem, *smem_k = smem + 2 * kFTileBytes, *smem_v = smem + 4 * kFTileBytes;
    uint64_t *bars = (uint64_t *)(smem + 6 * (size_t)kFTileBytes);
    uint64_t *qfull = bars, *kfull = bars + 1, *kempty = bars + 3, *vfull = bars + 5, *vempty = bars + 7;
    uint64_t *sfull = bars + 9, *pfull = bars + 11, *pvdone = bars + 13;       // [2]: query tile A, B
    uint32_t *tmem_slot = (uint32_t *)(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qp = (p.Sq + 2 * kFT - 1) / (2 * kFT);
    const int qp = n_qp - 1 - (int)(blockIdx.x % n_qp);          // heaviest (last) query pairs first
    const int h = (int)((blockIdx.x / n_qp) % p.H), b = (int)(blockIdx.x / ((size_t)n_qp * p.H));
    const int hk = h / (p.H / p.Hkv);
    const int kv_tiles = (p.Skv + kFT - 1) / kFT;
    int TX[2];                                                   // key tiles each query tile sees
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const int qt = 2 * qp + x;
        TX[x] = qt * kFT < p.Sq ? ((p.causal && qt + 1 < kv_tiles) ? qt + 1 : kv_tiles) : 0;
    }
    const int T = TX[0] > TX[1] ? TX[0] : TX[1];

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(qfull), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1);
            mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1);
            mbar_init(smem_u32(&sfull[s]), 1); mbar_init(smem_u32(&pfull[s]), 4); mbar_init(smem_u32(&pvdone[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // all 512 TMEM columns: S_A [0,128), S_B [128,256), O_A [256,384), O_B [384,512); P_x = first 64 columns of S_x
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t qf = smem_u32(qfull);
            mbar_expect_tx(qf, 2 * kFTileBytes);
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t qd = smem_u32(smem_q + (size_t)x * kFTileBytes);
                tma_load_4d(qd, &map_q, 0, (2 * qp + x) * kFT, h, b, qf);
                tma_load_4d(qd + kFHalf, &map_q, 64, (2 * qp + x) * kFT, h, b, qf);
            }
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                const uint32_t par = (uint32_t)(((t >> 1) & 1) ^ 1);
                mbar_wait(smem_u32(&kempty[s]), par);
                const uint32_t kf = smem_u32(&kfull[s]);
                mbar_expect_tx(kf, kFTileBytes);
                const uint32_t kd = smem_u32(smem_k + (size_t)s * kFTileBytes);
                tma_load_4d(kd, &map_k, 0, t * kFT, hk, b, kf);
                tma_load_4d(kd + kFHalf, &map_k, 64, t * kFT, hk, b, kf);
                mbar_wait(smem_u32(&vempty[s]), par);
                const uint32_t vf = smem_u32(&vfull[s]);
                mbar_expect_tx(vf, kFTileBytes);
                const uint32_t vd = smem_u32(smem_v + (size_t)s * kFTileBytes);
                tma_load_4d(vd, &map_v, 0, t * kFT, hk, b, vf);
                tma_load_4d(vd + kFHalf, &map_v, 64, t * kFT, hk, b, vf);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread).  The tensor pipe executes in issue order, so Q_x K(t+1)^T (which overwrites S_x, hence
        // P_x) may be issued right behind P_x(t) V(t) without any barrier in between. =====
        if (lane == 0) {
            auto issue_qk = [&](int x, int t) {
                const uint32_t qbase = smem_u32(smem_q + (size_t)x * kFTileBytes);
                const uint32_t kbase = smem_u32(smem_k + (size_t)(t & 1) * kFTileBytes);
                const uint32_t dcol = tmem_base + (uint32_t)(x * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = (uint32_t)(kq * kFHalf + kk * 32);
                        umma_f16(dcol, umma_desc_sw128(qbase + off), umma_desc_sw128(kbase + off), p.idesc_qk, (uint32_t)((kq | kk) != 0));
                    }
                umma_commit(smem_u32(&sfull[x]));
            };
            auto issue_pv = [&](int x, int t) {
                const uint32_t vbase = smem_u32(smem_v + (size_t)(t & 1) * kFTileBytes);
                const uint32_t pcol = tmem_base + (uint32_t)(x * kFT);              // P_x: 64 columns, 8 per 16 keys
                const uint32_t ocol = tmem_base + (uint32_t)(2 * kFT + x * kFT);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                                   // 16 keys per MMA
                    const uint32_t voff = (uint32_t)(ks * 16 * 128);               // 16 key rows of 128 B in each head-dim half
                    umma_f16_ts(ocol, pcol + (uint32_t)(ks * 8), umma_desc_sw128_mn(vbase + voff, kFHalf), p.idesc_pv,
                                (uint32_t)((t | ks) != 0));
                }
                umma_commit(smem_u32(&pvdone[x]));
            };
            mbar_wait(smem_u32(qfull), 0);
            mbar_wait(smem_u32(&kfull[0]), 0);
            tc_fence_after();
            if (TX[0] > 0) issue_qk(0, 0);
            if (TX[1] > 0) issue_qk(1, 0);
            umma_commit(smem_u32(&kempty[0]));
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                bool k_next_ready = false, v_ready = false;
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    if (t < TX[x]) {
                        mbar_wait(smem_u32(&pfull[x]), (uint32_t)(t & 1));           // P_x(t) is in TMEM, O_x rescaled
                        if (!v_ready) { mbar_wait(smem_u32(&vfull[s]), (uint32_t)((t >> 1) & 1)); v_ready = true; }
                        tc_fence_after();
                        issue_pv(x, t);
                    }
                    if (t + 1 < TX[x]) {
                        if (!k_next_ready) { mbar_wait(smem_u32(&kfull[(t + 1) & 1]), (uint32_t)(((t + 1) >> 1) & 1)); k_next_ready = true; }
                        tc_fence_after();
                        issue_qk(x, t + 1);
                    }
                }
                umma_commit(smem_u32(&vempty[s]));
                if (t + 1 < T) umma_commit(smem_u32(&kempty[(t + 1) & 1]));
            }
        }
    } else {
        // ===== softmax + epilogue: warpgroup x = query tile x; thread = one query row, all 128 key columns of a tile =====
        const int quad = warp & 3;
        const int x = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const int qrow0 = (2 * qp + x) * kFT;
        const int i = qrow0 + row;                             // query index
        const int Tx = TX[x];
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t s_col = lane_base + (uint32_t)(x * kFT), o_col = lane_base + (uint32_t)(2 * kFT + x * kFT);
        float m_used = -INFINITY, l = 0.0f, mass_acc = 0.0f;
        for (int t = 0; t < Tx; ++t) {
            mbar_wait(smem_u32(&sfull[x]), (uint32_t)(t & 1));
            tc_fence_after();
            uint32_t raw[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32_issue(s_col + (uint32_t)(32 * c), raw[c]);
            tmem_ld_wait();
            const int j0 = t * kFT;
            const bool need_mask = (p.causal && j0 + kFT - 1 > qrow0) || (j0 + kFT > p.Skv);
            float mx = -INFINITY;
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = j0 + 32 * c + e;
                        const bool dead = j >= p.Skv || (p.causal && j > i);
                        mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(raw[c][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(raw[c][e]));
            }
            mx *= p.scale_log2;                                  // log2 domain from here
            float alpha = 1.0f;
            bool rescale = false;
            if (mx > m_used + 8.0f) {                            // lazy rescale (see the single-tile kernel)
                rescale = m_used != -INFINITY;
                alpha = rescale ? ex2_approx(m_used - mx) : 0.0f;
                m_used = mx;
            }
            if (__any_sync(0xffffffffu, rescale)) {
                mbar_wait(smem_u32(&pvdone[x]), (uint32_t)((t - 1) & 1));          // P_x(t-1) V(t-1) has retired: O_x stable
                tc_fence_after();
                const float a = rescale ? alpha : 1.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(o_col + (uint32_t)(32 * c), o);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a);
                    tmem_st32(o_col + (uint32_t)(32 * c), o);
                }
                tmem_st_wait();
            }
            if (rescale) { l *= alpha; mass_acc *= alpha; }
            // p = 2^(s*scale - m) -> packed pairs written over the first 64 columns of S_x
            float rs = 0.0f, ps = 0.0f;
            const bool slow = need_mask || t == 0;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t pk[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const int c = 2 * c2 + (u >> 4), e = 2 * (u & 15);
                    float p0 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e]), p.scale_log2, -m_used));
                    float p1 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e + 1]), p.scale_log2, -m_used));
                    if (slow) {
                        const int j = j0 + 32 * c + e;
                        if (j >= p.Skv || (p.causal && j > i)) p0 = 0.0f;
                        if (j + 1 >= p.Skv || (p.causal && j + 1 > i)) p1 = 0.0f;
                        if (t == 0) ps += (j < p.P ? p0 : 0.0f) + (j + 1 < p.P ? p1 : 0.0f);
                    }
                    rs += p0 + p1;
                    pk[u] = pack2<BF16>(p0, p1);
                }
                tmem_st32(s_col + (uint32_t)(32 * c2), pk);
            }
            tmem_st_wait();
            l += rs;
            if (t == 0) mass_acc = ps;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&pfull[x]));
        }
        if (Tx > 0) {
            mbar_wait(smem_u32(&pvdone[x]), (uint32_t)((Tx - 1) & 1));
            tc_fence_after();
            const float inv_l = 1.0f / l;
            const bool live = i < p.Sq;
            unsigned char *orow = (unsigned char *)p.out + ((size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)i * p.o_ss) * 2;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t o[32];
                tmem_ld32(o_col + (uint32_t)(32 * c), o);
                if (live) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = pack2<BF16>(__uint_as_float(o[8 * q + 2 * e]) * inv_l, __uint_as_float(o[8 * q + 2 * e + 1]) * inv_l);
                        *reinterpret_cast<uint4 *>(orow + (size_t)(32 * c + 8 * q) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            if (live) {
                const size_t at = ((size_t)b * p.H + h) * p.Sq + i;
                if (p.lse) p.lse[at] = (m_used + log2f(l)) * 0.6931471805599453f;
                if (p.head_sums) p.head_sums[at] = mass_acc * inv_l;
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- alternating key tiles (rtkv_debug_flash_variant(3)) ------------------------------------------------------------------
// The two kernels above keep both softmax warps of a scheduler on the SAME key tile, so they reach the exponentials together:
// the MUFU pipe (8 cycles per warp instruction) is saturated for 1024 cycles per tile and idle for the other ~1500 (TMEM
// load, row max, packs, shared-memory stores).  Here warpgroup g owns the key tiles t = g, g + 2, ... outright (a thread = one
// query row x all 128 key columns of the tile, held in registers: no exchange between threads), with its own S buffer, P
// buffer, running max / sum and accumulator O_g in tensor memory.  Q K_{t+2}^T is issued as soon as warpgroup g has S_t in
// registers, so a warpgroup never waits for the tensor pipe and the two warpgroups drift half a tile apart: one runs its
// exponentials while the other loads, maxes and stores.  The two partial results (even tiles, odd tiles) are merged in the
// epilogue exactly like the independent-halves kernel: O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_g = 2^(m_g - M).
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr int kATThreads = 256 + 128;           // two softmax warpgroups + one warpgroup holding the TMA and MMA warps
template <bool BF16, bool PROF = false>
__global__ void __launch_bounds__(kATThreads, 1)
flash_prefill_at_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, FlashParams p)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_q = smem, *smem_k = smem + kFTileBytes, *smem_v = smem + 3 * kFTileBytes, *smem_p = smem + 5 * kFTileBytes;
    uint64_t *bars = (uint64_t *)(smem + 7 * (size_t)kFTileBytes);
    uint64_t *qfull = bars, *kfull = bars + 1, *kempty = bars + 3, *vfull = bars + 5, *vempty = bars + 7;
    uint64_t *sfull = bars + 9, *sfree = bars + 11, *pfull = bars + 13, *pvdone = bars + 15;   // [2]: tile parity = warpgroup
    uint32_t *tmem_slot = (uint32_t *)(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (p.Sq + kFT - 1) / kFT;
    const int qt = n_qt - 1 - (int)(blockIdx.x % n_qt);          // heaviest (last) query tiles first
    const int h = (int)((blockIdx.x / n_qt) % p.H), b = (int)(blockIdx.x / ((size_t)n_qt * p.H));
    const int hk = h / (p.H / p.Hkv);
    int T = (p.Skv + kFT - 1) / kFT;                             // key tiles this query tile sees
    if (p.causal && qt + 1 < T) T = qt + 1;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(qfull), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1);
            mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1);
            mbar_init(smem_u32(&sfull[s]), 1); mbar_init(smem_u32(&sfree[s]), 4);
            mbar_init(smem_u32(&pfull[s]), 4); mbar_init(smem_u32(&pvdone[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {   // 512 TMEM columns: S_0 [0,128), S_1 [128,256), O_0 [256,384), O_1 [384,512) (index = tile parity)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    // warps 0-3 / 4-7: softmax warpgroups (128 logits of a row live in registers); warps 8-11: TMA, MMA, two idle.  Per scheduler
    // 2 x 216 + 72 = 504 = 3 x 168 registers per lane: what the CTA was given at launch (the pool setmaxnreg trades within).
    // (setmaxnreg sits at the top of each role's branch so that ptxas sees which budget the code below it runs under)
    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 8) {
        // ===== TMA producer, Q and K.  K and V have separate producer warps: in one loop the wait for V's slot (P_{t-2} V_{t-2}
        // retired) also held back the next K tile, and Q K_{t+2}^T is wanted long before that =====
        if (lane == 0) {
            const uint32_t qf = smem_u32(qfull);
            mbar_expect_tx(qf, kFTileBytes);
            tma_load_4d(smem_u32(smem_q), &map_q, 0, qt * kFT, h, b, qf);
            tma_load_4d(smem_u32(smem_q) + kFHalf, &map_q, 64, qt * kFT, h, b, qf);
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                mbar_wait(smem_u32(&kempty[s]), (uint32_t)(((t >> 1) & 1) ^ 1));
                const uint32_t kf = smem_u32(&kfull[s]);
                mbar_expect_tx(kf, kFTileBytes);
                const uint32_t kd = smem_u32(smem_k + (size_t)s * kFTileBytes);
                tma_load_4d(kd, &map_k, 0, t * kFT, hk, b, kf);
                tma_load_4d(kd + kFHalf, &map_k, 64, t * kFT, hk, b, kf);
            }
        }
    } else if (warp == 10) {
        // ===== TMA producer, V =====
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                mbar_wait(smem_u32(&vempty[s]), (uint32_t)(((t >> 1) & 1) ^ 1));
                const uint32_t vf = smem_u32(&vfull[s]);
                mbar_expect_tx(vf, kFTileBytes);
                const uint32_t vd = smem_u32(smem_v + (size_t)s * kFTileBytes);
                tma_load_4d(vd, &map_v, 0, t * kFT, hk, b, vf);
                tma_load_4d(vd + kFHalf, &map_v, 64, t * kFT, hk, b, vf);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer (one thread): Q K_0, Q K_1, then per tile t: Q K_{t+2} (its S buffer is free as soon as the owning
        // warpgroup holds S_t in registers), P_t V_t into O_{t&1} =====
        if (lane == 0) {
            const uint32_t qbase = smem_u32(smem_q);
            auto issue_qk = [&](int t) {
                const int s = t & 1;
                mbar_wait(smem_u32(&kfull[s]), (uint32_t)((t >> 1) & 1));
                mbar_wait(smem_u32(&sfree[s]), (uint32_t)(((t >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t kbase = smem_u32(smem_k + (size_t)s * kFTileBytes);
                const uint32_t dcol = tmem_base + (uint32_t)(s * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = (uint32_t)(kq * kFHalf + kk * 32);
                        umma_f16(dcol, umma_desc_sw128(qbase + off), umma_desc_sw128(kbase + off), p.idesc_qk, (uint32_t)((kq | kk) != 0));
                    }
                umma_commit(smem_u32(&kempty[s]));
                umma_commit(smem_u32(&sfull[s]));
            };
            mbar_wait(smem_u32(qfull), 0);
            issue_qk(0);
            if (T > 1) issue_qk(1);
            long long mma_wait_p = 0, mma_wait_qk = 0, mma_t0 = 0;
            if constexpr (PROF) mma_t0 = clock64();
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                long long w0 = 0;
                if constexpr (PROF) w0 = clock64();
                if (t + 2 < T) issue_qk(t + 2);
                if constexpr (PROF) { const long long w1 = clock64(); mma_wait_qk += w1 - w0; w0 = w1; }
                mbar_wait(smem_u32(&pfull[s]), (uint32_t)((t >> 1) & 1));            // P_t in shared memory, O_s rescaled
                if constexpr (PROF) mma_wait_p += clock64() - w0;
                mbar_wait(smem_u32(&vfull[s]), (uint32_t)((t >> 1) & 1));
                tc_fence_after();
                const uint32_t vbase = smem_u32(smem_v + (size_t)s * kFTileBytes);
                const uint32_t pbase = smem_u32(smem_p + (size_t)s * kFTileBytes);
                const uint32_t ocol = tmem_base + (uint32_t)(2 * kFT + s * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t poff = (uint32_t)(kq * kFHalf + kk * 32);
                        const uint32_t voff = (uint32_t)((kq * 64 + kk * 16) * 128);
                        umma_f16(ocol, umma_desc_sw128(pbase + poff), umma_desc_sw128_mn(vbase + voff, kFHalf), p.idesc_pv,
                                 (uint32_t)(((t >> 1) | kq | kk) != 0));
                    }
                umma_commit(smem_u32(&vempty[s]));
                umma_commit(smem_u32(&pvdone[s]));
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0 && p.prof) { p.prof[32] = mma_wait_p; p.prof[33] = clock64() - mma_t0; p.prof[34] = mma_wait_qk; p.prof[35] = T; }
            }
        }
    }
    } else {
        // ===== softmax + epilogue: warpgroup g = key tiles of parity g; thread = one query row x 128 key columns =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        const int quad = warp & 3;
        const int g = warp >> 2;
        const int row = quad * 32 + lane;
        const int i = qt * kFT + row;                          // query index
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t s_col = lane_base + (uint32_t)(g * kFT), o_col = lane_base + (uint32_t)(2 * kFT + g * kFT);
        const uint32_t sfull_g = smem_u32(&sfull[g]), sfree_g = smem_u32(&sfree[g]), pfull_g = smem_u32(&pfull[g]), pvdone_g = smem_u32(&pvdone[g]);
        float m_used = -INFINITY, l = 0.0f, mass_acc = 0.0f;   // this warpgroup's tiles only
        long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_prev = 0;
        auto tick = [&](int k) {
            if constexpr (PROF) {
                const long long now = clock64();
                prof_acc[k] += now - prof_prev;
                prof_prev = now;
            }
        };
        if constexpr (PROF) prof_prev = clock64();
        for (int t = g; t < T; t += 2) {
            const uint32_t par = (uint32_t)((t >> 1) & 1);
            tick(0);
            mbar_wait(sfull_g, par);
            tc_fence_after();
            tick(1);
            uint32_t raw[2][2][32];                              // [64-column half][32-column chunk][column]
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32_issue(s_col + (uint32_t)(32 * c), raw[c >> 1][c & 1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sfree_g);
            tick(2);
            const int j0 = t * kFT;
            const bool need_mask = (p.causal && j0 + kFT - 1 > qt * kFT) || (j0 + kFT > p.Skv);
            float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = j0 + 32 * c + e;
                        const bool dead = j >= p.Skv || (p.causal && j > i);
                        mq[c] = fmaxf(mq[c], dead ? -INFINITY : __uint_as_float(raw[c >> 1][c & 1][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 32; e += 2)
                        mq[c] = fmax3(mq[c], __uint_as_float(raw[c >> 1][c & 1][e]), __uint_as_float(raw[c >> 1][c & 1][e + 1]));
            }
            const float mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])) * p.scale_log2;   // log2 domain from here
            float alpha = 1.0f;
            bool rescale = false;
            if (mx > m_used + 8.0f) {                            // lazy rescale (see flash_prefill_kernel)
                rescale = m_used != -INFINITY;
                alpha = rescale ? ex2_approx(m_used - mx) : 0.0f;
                m_used = mx;
            }
            if constexpr (PROF) { if (mx == 12345.678f) prof_acc[7] += 1; }
            tick(3);
            mbar_wait(pvdone_g, par ^ 1u);                       // P_{t-2} V_{t-2} has retired: O_g stable, P buffer g free
            tick(5);
            if (__any_sync(0xffffffffu, rescale)) {              // warp-uniform: tcgen05.ld/st are warp-wide; rare
                tc_fence_after();
                const float a = rescale ? alpha : 1.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(o_col + (uint32_t)(32 * c), o);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a);
                    tmem_st32(o_col + (uint32_t)(32 * c), o);
                }
                tmem_st_wait();
            }
            if (rescale) { l *= alpha; mass_acc *= alpha; }
            float rs = 0.0f, ps = 0.0f;
            const uint32_t prow = smem_u32(smem_p) + (uint32_t)(g * kFTileBytes + row * 128);
            if (need_mask || t == 0) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    exp_store<BF16, true, 0>(raw[hf], p.scale_log2, m_used, prow + (uint32_t)(hf * kFHalf), row, j0 + 64 * hf, i, p.Skv, p.causal, p.P, t == 0, rs, ps);
            } else {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    exp_store<BF16, false, 0>(raw[hf], p.scale_log2, m_used, prow + (uint32_t)(hf * kFHalf), row, j0 + 64 * hf, i, p.Skv, p.causal, p.P, false, rs, ps);
            }
            l += rs;
            if (t == 0) mass_acc = ps;
            if constexpr (PROF) { if (l == 12345.678f) prof_acc[7] += 1; }
            tick(6);
            fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull_g);
            tick(7);
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && lane == 0 && quad == 0 && p.prof) {
                for (int k = 0; k < 8; ++k) p.prof[g * 16 + k] = prof_acc[k];
                p.prof[g * 16 + 8] = (T - g + 1) >> 1;
            }
        }
        // ---- epilogue: merge the even-tile and odd-tile online softmaxes.  Both accumulators have retired first, so the P
        // buffers are free: they carry the three floats per (row, warpgroup).
        const int n0 = (T + 1) >> 1, n1 = T >> 1;              // tiles of parity 0 / 1
        // Each warpgroup waits for ITS accumulator only (a parity wait is only meaningful on a barrier whose phases the
        // waiter has followed: the other warpgroup's may be several phases behind or ahead), then the two meet.
        const int n_g = g == 0 ? n0 : n1;
        if (n_g > 0) mbar_wait(pvdone_g, (uint32_t)((n_g - 1) & 1));
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        tc_fence_after();
        float *xe = reinterpret_cast<float *>(smem_p);           // [3 values][2 warpgroups][128 rows]
        xe[(0 * 2 + g) * kFT + row] = m_used;
        xe[(1 * 2 + g) * kFT + row] = l;
        xe[(2 * 2 + g) * kFT + row] = mass_acc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float m_o = xe[(0 * 2 + (g ^ 1)) * kFT + row], l_o = xe[(1 * 2 + (g ^ 1)) * kFT + row], mass_o = xe[(2 * 2 + (g ^ 1)) * kFT + row];
        const float M = fmaxf(m_used, m_o);
        const float a_me = ex2_approx(m_used - M), a_ot = ex2_approx(m_o - M);       // 2^(-inf) = 0 for a warpgroup that saw no tile
        const float L = a_me * l + a_ot * l_o;
        const float inv_l = 1.0f / L;
        const float a0 = (g == 0 ? a_me : a_ot) * inv_l, a1 = (g == 0 ? a_ot : a_me) * inv_l;
        const bool live = i < p.Sq;
        unsigned char *orow = (unsigned char *)p.out + ((size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)i * p.o_ss) * 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                            // this thread writes output columns [64g, 64g + 64)
            uint32_t o0[32], o1[32];
            tmem_ld32_issue(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o0);
            if (n1 > 0) tmem_ld32_issue(lane_base + (uint32_t)(3 * kFT + 64 * g + 32 * c), o1);   // O_1 was never written when T == 1
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c0 = 8 * q + 2 * e;
                        float x0 = __uint_as_float(o0[c0]) * a0, x1 = __uint_as_float(o0[c0 + 1]) * a0;
                        if (n1 > 0) { x0 += __uint_as_float(o1[c0]) * a1; x1 += __uint_as_float(o1[c0 + 1]) * a1; }
                        w[e] = pack2<BF16>(x0, x1);
                    }
                    *reinterpret_cast<uint4 *>(orow + (size_t)(64 * g + 32 * c + 8 * q) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        if (live && g == 0) {
            const size_t at = ((size_t)b * p.H + h) * p.Sq + i;
            if (p.lse) p.lse[at] = (M + log2f(L)) * 0.6931471805599453f;
            if (p.head_sums) p.head_sums[at] = (a_me * mass_acc + a_ot * mass_o) * inv_l;
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- alternating key tiles, ONE accumulator, P in tensor memory (rtkv_debug_flash_variant(4)) -----------------------------
// What the phase profile (tools/profile_flash_phases.py) of the kernels above showed: (1) both softmax warps of a scheduler reach
// the exponentials together, so MUFU is saturated for ~1300 cycles of a tile and idle for the other ~1100; (2) an SS-mode
// 128x128x16 tcgen05.mma reads 8 KB of operands from shared memory in its 64 cycles = the whole 128 B/clk of the SM, so with the
// TMA writes (64 KB per tile) and the P stores (32 KB) a tile moves 224 KB through shared memory = 1750 cycles against 1024 of
// tensor work; (3) with two K stages a K tile is only requested one tile time before it is needed, less than the TMA latency.
// This kernel: warpgroup g owns the key tiles t = g, g + 2, ... (a thread = one query row x 128 key columns, in registers), so
// the two warps of a scheduler are half a tile apart (one in MUFU, one in load / max / store).  The online softmax stays ONE
// chain: the owner of tile t takes the reference max m_{t-1} from the owner of tile t-1 (one float per row through shared
// memory + a named producer/consumer barrier), publishes m_t right after its row max -- before its exponentials -- and the
// single accumulator O is rescaled (rarely: lazy threshold 2^8) by whoever raises the reference, once P_{t-1} V_{t-1} has
// retired.  P goes to TENSOR memory (two bf16/f16 per 32-bit column, 64 columns per buffer) and is the A operand of the second
// product from there (tcgen05.mma with A in TMEM): no P traffic in shared memory (160 KB per tile), which makes room for THREE
// K and V stages.  TMEM: S_0 [0,128) S_1 [128,256) O [256,384) P_0 [384,448) P_1 [448,512).
constexpr int kMCStages = 3;
constexpr size_t kMCSmem = 768 + (1 + 2 * kMCStages) * (size_t)kFTileBytes + 256 + 2 * kFT * 8;   // = 232448: all a CTA can have

// warp-convergent code only: true in exactly one lane
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// reference-max hand-off between the two softmax warpgroups: {value, tile} in one aligned 8-byte shared-memory word
__device__ __forceinline__ void xm_put(uint32_t addr, float m, int tile)
{
    asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(m)), "r"(tile) : "memory");
}
__device__ __forceinline__ float xm_take(uint32_t addr, int tile)
{
    uint32_t v, tag;
    const long long t0 = clock64();
    do {
        asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(tag) : "r"(addr) : "memory");
        if (tag != (uint32_t)tile && clock64() - t0 > 4000000000LL) __trap();
    } while (tag != (uint32_t)tile);
    return __uint_as_float(v);
}

template <bool BF16, bool SLOW, int PE = 0>
__device__ __forceinline__ void exp_pack(const uint32_t (&raw)[2][32], float scale_log2, float m_used, uint32_t (&pk)[32],
                                         int j_base, int i, int Skv, int causal, int P, bool first, float &rs, float &ps)
{
    if constexpr (!SLOW && PE == 0) {
        // The unmasked tile: four independent partial row sums instead of one serial chain of adds (tools/flash_timeline.py: a
        // warp with the MUFU pipe to itself needed 1850 cycles per 128-column row, 14 per exponential against the pipe's 8).
        // ptxas interleaves packs and adds two exponentials behind their MUFU whatever the source order; with both warps of a
        // scheduler in this phase the pipe is saturated (2300 cycles for 2 x 128 exponentials), alone a warp reaches about half.
        float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float pv[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) pv[e] = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e]), scale_log2, -m_used));
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                r0 += pv[e]; r1 += pv[e + 1]; r2 += pv[e + 2]; r3 += pv[e + 3];
                pk[16 * c + (e >> 1)] = pack2<BF16>(pv[e], pv[e + 1]);
                pk[16 * c + (e >> 1) + 1] = pack2<BF16>(pv[e + 2], pv[e + 3]);
            }
        }
        rs += (r0 + r1) + (r2 + r3);
        return;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
            float p0, p1;
            if (!SLOW && PE > 0 && ((e >> 1) % (PE > 0 ? PE : 1)) == PE - 1) {     // this pair on the FMA pipe
                p0 = exp2_fma(__fmaf_rn(__uint_as_float(raw[c][e]), scale_log2, -m_used));
                p1 = exp2_fma(__fmaf_rn(__uint_as_float(raw[c][e + 1]), scale_log2, -m_used));
            } else {
                p0 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e]), scale_log2, -m_used));
                p1 = ex2_approx(__fmaf_rn(__uint_as_float(raw[c][e + 1]), scale_log2, -m_used));
            }
            if constexpr (SLOW) {
                const int j = j_base + 32 * c + e;
                if (j >= Skv || (causal && j > i)) p0 = 0.0f;
                if (j + 1 >= Skv || (causal && j + 1 > i)) p1 = 0.0f;
                if (first) ps += (j < P ? p0 : 0.0f) + (j + 1 < P ? p1 : 0.0f);
            }
            rs += p0 + p1;
            pk[16 * c + (e >> 1)] = pack2<BF16>(p0, p1);
        }
}

template <bool BF16, bool PROF = false, int PE = 0>
__global__ void __launch_bounds__(kATThreads, 1)
flash_prefill_mc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, FlashParams p)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_q = smem, *smem_k = smem + kFTileBytes, *smem_v = smem + (1 + kMCStages) * (size_t)kFTileBytes;
    uint64_t *bars = (uint64_t *)(smem + (1 + 2 * kMCStages) * (size_t)kFTileBytes);
    uint64_t *qfull = bars, *kfull = bars + 1, *kempty = bars + 4, *vfull = bars + 7, *vempty = bars + 10;
    uint64_t *sfull = bars + 13, *sfree = bars + 15, *pfull = bars + 17, *pvdone = bars + 19;   // [2]: tile parity = warpgroup
    uint32_t *tmem_slot = (uint32_t *)(bars + 21);
    uint32_t *xm = (uint32_t *)((unsigned char *)bars + 256);    // [2 tile parities][128 rows] x {reference max after tile t, t}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = (p.Sq + kFT - 1) / kFT;
    const int qt = n_qt - 1 - (int)(blockIdx.x % n_qt);          // heaviest (last) query tiles first
    const int h = (int)((blockIdx.x / n_qt) % p.H), b = (int)(blockIdx.x / ((size_t)n_qt * p.H));
    const int hk = h / (p.H / p.Hkv);
    int T = (p.Skv + kFT - 1) / kFT;                             // key tiles this query tile sees
    if (p.causal && qt + 1 < T) T = qt + 1;

    if (threadIdx.x < 2 * kFT) xm[2 * threadIdx.x + 1] = 0xffffffffu;   // no tile published yet
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(qfull), 1);
        for (int s = 0; s < kMCStages; ++s) {
            mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1);
            mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&sfull[s]), 1); mbar_init(smem_u32(&sfree[s]), 4);
            mbar_init(smem_u32(&pfull[s]), 4); mbar_init(smem_u32(&pvdone[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    // PROF: per-tile cycle stamps (tools/flash_timeline.py) are on when counters[62] != 0; read once, not per stamp
    bool stamps = false;
    if constexpr (PROF) stamps = blockIdx.x == 0 && p.prof != nullptr && p.prof[62] != 0;
    // warps 0-3 / 4-7: softmax warpgroups; warps 8-11: Q+K producer, Q K^T issuer, V producer, P V issuer.  2 x 216 + 72 = 504 = 3 x 168
    // registers per lane: what the CTA was given at launch (setmaxnreg trades within that pool; it sits at the top of each role's
    // branch so that ptxas sees which budget the code below it runs under).
    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 8) {
        // ===== TMA producer: Q, then the K ring =====
        if (lane == 0) {
            const uint32_t qf = smem_u32(qfull);
            mbar_expect_tx(qf, kFTileBytes);
            tma_load_4d(smem_u32(smem_q), &map_q, 0, qt * kFT, h, b, qf);
            tma_load_4d(smem_u32(smem_q) + kFHalf, &map_q, 64, qt * kFT, h, b, qf);
            int s = 0;
            uint32_t ph = 0;
            long long w_empty = 0, w_lat = 0;                    // PROF: waiting for a free slot / issue -> data landed
            for (int t = 0; t < T; ++t) {
                long long c0 = 0;
                if constexpr (PROF) c0 = clock64();
                mbar_wait(smem_u32(&kempty[s]), ph ^ 1u);
                const uint32_t kf = smem_u32(&kfull[s]);
                mbar_expect_tx(kf, kFTileBytes);
                const uint32_t kd = smem_u32(smem_k + (size_t)s * kFTileBytes);
                long long c1 = 0;
                if constexpr (PROF) c1 = clock64();
                tma_load_4d(kd, &map_k, 0, t * kFT, hk, b, kf);
                tma_load_4d(kd + kFHalf, &map_k, 64, t * kFT, hk, b, kf);
                if constexpr (PROF) {                            // (serialises the K loads: only for the latency figure)
                    if (p.prof && p.prof[63] == 1) {
                        mbar_wait(kf, ph);
                        w_lat += clock64() - c1;
                    }
                    w_empty += c1 - c0;
                }
                if (++s == kMCStages) { s = 0; ph ^= 1u; }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0 && p.prof) { p.prof[44] = w_empty; p.prof[45] = w_lat; }
            }
        }
    } else if (warp == 10) {
        // ===== TMA producer: the V ring (its own warp: a V slot frees a whole tile later than a K slot) =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < T; ++t) {
                mbar_wait(smem_u32(&vempty[s]), ph ^ 1u);
                const uint32_t vf = smem_u32(&vfull[s]);
                mbar_expect_tx(vf, kFTileBytes);
                const uint32_t vd = smem_u32(smem_v + (size_t)s * kFTileBytes);
                tma_load_4d(vd, &map_v, 0, t * kFT, hk, b, vf);
                tma_load_4d(vd + kFHalf, &map_v, 64, t * kFT, hk, b, vf);
                if (++s == kMCStages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer 1 (one thread): the first products, Q K_t^T -> S_{t&1}.  The two products have separate issuing
        // threads: with one, issuing (~80 cycles per MMA) plus its four waits per tile made the issuing thread, not the tensor
        // pipe, the slowest stage (tools/profile_flash_phases.py).  Nothing orders the two streams against each other but the
        // mbarriers they already share with the softmax warps: Q K_{t+2} only needs K_{t+2} and S_t in its owner's registers.
        // Descriptors are built once and advanced by one add per MMA (umma.cuh). =====
        if (lane == 0) {
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem_q)), kd0 = umma_desc_sw128(smem_u32(smem_k));
            const uint32_t idesc_qk = p.idesc_qk;
            int ks_ = 0;                                         // K ring position
            uint32_t kph = 0;
            long long mw[3] = {0, 0, 0}, mprev = 0;              // PROF: K wait, S-buffer wait, QK issue
            auto mtick = [&](int k) {
                if constexpr (PROF) {
                    const long long now = clock64();
                    mw[k] += now - mprev;
                    mprev = now;
                }
            };
            mbar_wait(smem_u32(qfull), 0);
            long long mma_t0 = 0;
            if constexpr (PROF) { mma_t0 = clock64(); mprev = mma_t0; }
            for (int t = 0; t < T; ++t) {
                const int sb = t & 1;
                mbar_wait(smem_u32(&kfull[ks_]), kph);
                mtick(0);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 0] = clock64(); }
                mbar_wait(smem_u32(&sfree[sb]), (uint32_t)(((t >> 1) & 1) ^ 1));     // S_{t-2} is in its owner's registers
                mtick(1);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 1] = clock64(); }
                tc_fence_after();
                const uint64_t kd = umma_desc_advance(kd0, (uint32_t)(ks_ * kFTileBytes));
                const uint32_t dcol = tmem_base + (uint32_t)(sb * kFT);
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = (uint32_t)(kq * kFHalf + kk * 32);
                        umma_f16(dcol, umma_desc_advance(qd0, off), umma_desc_advance(kd, off), idesc_qk, (uint32_t)((kq | kk) != 0));
                    }
                umma_commit(smem_u32(&kempty[ks_]));
                umma_commit(smem_u32(&sfull[sb]));
                if (++ks_ == kMCStages) { ks_ = 0; kph ^= 1u; }
                mtick(2);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 2] = clock64(); }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0 && p.prof) {
                    p.prof[34] = clock64() - mma_t0; p.prof[35] = T;
                    for (int k = 0; k < 3; ++k) p.prof[36 + k] = mw[k];
                }
            }
        }
    } else {
        // ===== MMA issuer 2 (warp 11, one thread): the second products, O += P_t V_t with P_t read from tensor memory =====
        if (lane == 0) {
            const uint64_t vd0 = umma_desc_sw128_mn(smem_u32(smem_v), kFHalf);
            const uint32_t idesc_pv = p.idesc_pv;
            long long mw[3] = {0, 0, 0}, mprev = 0;              // PROF: P wait, V wait, PV issue
            auto mtick = [&](int k) {
                if constexpr (PROF) {
                    const long long now = clock64();
                    mw[k] += now - mprev;
                    mprev = now;
                }
            };
            long long mma_t0 = 0;
            if constexpr (PROF) { mma_t0 = clock64(); mprev = mma_t0; }
            int vs_ = 0;
            uint32_t vph = 0;
            for (int t = 0; t < T; ++t) {
                const int sb = t & 1;
                mbar_wait(smem_u32(&pfull[sb]), (uint32_t)((t >> 1) & 1));           // P_t in tensor memory, O rescaled
                mtick(0);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 3] = clock64(); }
                mbar_wait(smem_u32(&vfull[vs_]), vph);
                mtick(1);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 4] = clock64(); }
                tc_fence_after();
                const uint64_t vd = umma_desc_advance(vd0, (uint32_t)(vs_ * kFTileBytes));
                const uint32_t pcol = tmem_base + (uint32_t)(3 * kFT + sb * 64);     // 8 columns per 16 keys
                const uint32_t ocol = tmem_base + (uint32_t)(2 * kFT);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                                     // 16 keys per MMA
                    const uint32_t voff = (uint32_t)(ks * 16 * 128);                 // 16 key rows of 128 B in each head-dim half
                    umma_f16_ts(ocol, pcol + (uint32_t)(ks * 8), umma_desc_advance(vd, voff), idesc_pv, (uint32_t)((t | ks) != 0));
                }
                umma_commit(smem_u32(&vempty[vs_]));
                umma_commit(smem_u32(&pvdone[sb]));
                if (++vs_ == kMCStages) { vs_ = 0; vph ^= 1u; }
                mtick(2);
                if constexpr (PROF) { if (stamps && t < 48) p.prof[384 + t * 6 + 5] = clock64(); }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0 && p.prof) {
                    p.prof[32] = mw[0]; p.prof[33] = clock64() - mma_t0;
                    for (int k = 0; k < 3; ++k) p.prof[39 + k] = mw[k];
                }
            }
        }
    }
    } else {
        // ===== softmax + epilogue: warpgroup g = key tiles of parity g; thread = one query row x 128 key columns =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        const int quad = warp & 3;
        const int g = warp >> 2;
        const int row = quad * 32 + lane;
        const int i = qt * kFT + row;                          // query index
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t s_col = lane_base + (uint32_t)(g * kFT), o_col = lane_base + (uint32_t)(2 * kFT);
        const uint32_t p_col = lane_base + (uint32_t)(3 * kFT + g * 64);
        const uint32_t sfull_g = smem_u32(&sfull[g]), sfree_g = smem_u32(&sfree[g]), pfull_g = smem_u32(&pfull[g]), pvdone_g = smem_u32(&pvdone[g]);
        const uint32_t pvdone_o = smem_u32(&pvdone[g ^ 1]);
        const uint32_t xm_u32 = smem_u32(xm);
        float m_own = -INFINITY;                                // the reference max this thread's l / mass_acc are expressed in
        float l = 0.0f, mass_acc = 0.0f;                        // this warpgroup's tiles only
        long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_prev = 0;
        auto tick = [&](int k) {
            if constexpr (PROF) {
                const long long now = clock64();
                prof_acc[k] += now - prof_prev;
                prof_prev = now;
            }
        };
        if constexpr (PROF) prof_prev = clock64();
        for (int t = g; t < T; t += 2) {
            const uint32_t par = (uint32_t)((t >> 1) & 1);
            tick(0);
            if constexpr (PROF) {
                if (stamps && lane == 0 && quad == 0 && (t >> 1) < 40) p.prof[64 + g * 128 + (t >> 1) * 3 + 2] = clock64();
            }
            mbar_wait(sfull_g, par);
            tc_fence_after();
            tick(1);
            uint32_t raw[2][2][32];                              // [64-column half][32-column chunk][column]
            const int j0 = t * kFT;
            const bool need_mask = (p.causal && j0 + kFT - 1 > qt * kFT) || (j0 + kFT > p.Skv);
            float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
            auto half_max = [&](int hf) {                        // row max over columns [64 hf, 64 hf + 64)
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int j = j0 + 64 * hf + 32 * c + e;
                            const bool dead = j >= p.Skv || (p.causal && j > i);
                            mq[2 * hf + c] = fmaxf(mq[2 * hf + c], dead ? -INFINITY : __uint_as_float(raw[hf][c][e]));
                        }
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int e = 0; e < 32; e += 2)
                            mq[2 * hf + c] = fmax3(mq[2 * hf + c], __uint_as_float(raw[hf][c][e]), __uint_as_float(raw[hf][c][e + 1]));
                }
            };
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32_issue(s_col + (uint32_t)(32 * c), raw[c >> 1][c & 1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sfree_g);
            tick(2);
            half_max(0);
            half_max(1);
            const float mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])) * p.scale_log2;   // log2 domain from here
            if constexpr (PROF) { if (mx == 12345.678f) prof_acc[7] += 1; }
            tick(3);
            // the reference max after the previous tile, from its owner (the other warpgroup): one 8-byte word {m, tile} per row
            // and tile parity, written and read by the two threads of the row -- no barrier, the reader spins on the tile tag
            float m_prev = -INFINITY;
            if (t > 0) m_prev = xm_take(xm_u32 + (uint32_t)((((t - 1) & 1) * kFT + row) * 8), t - 1);
            // lazy rescale: keep the old reference unless the row max grew by more than 2^8
            const bool bump = mx > m_prev + 8.0f;
            const bool rescale = bump && m_prev != -INFINITY;
            const float m_new = bump ? mx : m_prev;
            xm_put(xm_u32 + (uint32_t)(((t & 1) * kFT + row) * 8), m_new, t);       // published before anything slow
            if (m_new != m_own) {                                // bring this warpgroup's partial sums to the current reference
                const float a = ex2_approx(m_own - m_new);       // 2^(-inf) = 0 on the first own tile (l = mass_acc = 0 anyway)
                l *= a;
                mass_acc *= a;
                m_own = m_new;
            }
            tick(4);
            if (__any_sync(0xffffffffu, rescale)) {              // warp-uniform: tcgen05.ld/st are warp-wide; rare
                // O holds P_0 V_0 + ... + P_{t-1} V_{t-1} in units of m_prev once the previous tile's product has retired; the
                // other warpgroup follows that barrier, this one is never more than one phase away from it (P_t is still pending)
                mbar_wait(pvdone_o, (uint32_t)(((t - 1) >> 1) & 1));
                tc_fence_after();
                const float a = rescale ? ex2_approx(m_prev - m_new) : 1.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld32(o_col + (uint32_t)(32 * c), o);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a);
                    tmem_st32(o_col + (uint32_t)(32 * c), o);
                }
                tmem_st_wait();
            }
            mbar_wait(pvdone_g, par ^ 1u);                       // P_{t-2} V_{t-2} has retired: P buffer g is free
            tc_fence_after();
            tick(5);
            if constexpr (PROF) {
                if (stamps && lane == 0 && quad == 0 && (t >> 1) < 40) p.prof[64 + g * 128 + (t >> 1) * 3] = clock64();
            }
            // The hand-off above takes a few hundred cycles and the exponentials over a thousand, so left alone the two
            // warpgroups run their exponentials at the same time (sharing the MUFU pipe) and their loads / maxima at the same
            // time (MUFU idle).  Nothing in the chain forces that: the odd warpgroup starts half a tile late, once, and the
            // offset persists -- from then on one warp of each scheduler is in MUFU while the other loads, maxes and stores.
            if (t == 1 && T > 4 && p.skew > 0) {
                const long long c0 = clock64();
                while (clock64() - c0 < (long long)p.skew) { }
            }
            float rs = 0.0f, ps = 0.0f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t pk[32];
                if (need_mask || t == 0)
                    exp_pack<BF16, true>(raw[hf], p.scale_log2, m_new, pk, j0 + 64 * hf, i, p.Skv, p.causal, p.P, t == 0, rs, ps);
                else
                    exp_pack<BF16, false, PE>(raw[hf], p.scale_log2, m_new, pk, j0 + 64 * hf, i, p.Skv, p.causal, p.P, false, rs, ps);
                tmem_st32(p_col + (uint32_t)(32 * hf), pk);
            }
            l += rs;
            if (t == 0) mass_acc = ps;
            if constexpr (PROF) { if (l == 12345.678f) prof_acc[7] += 1; }
            tick(6);
            if constexpr (PROF) {
                if (stamps && lane == 0 && quad == 0 && (t >> 1) < 40) p.prof[64 + g * 128 + (t >> 1) * 3 + 1] = clock64();
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull_g);
            tick(7);
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && lane == 0 && quad == 0 && p.prof) {
                for (int k = 0; k < 8; ++k) p.prof[g * 16 + k] = prof_acc[k];
                p.prof[g * 16 + 8] = (T - g + 1) >> 1;
            }
        }
        // ---- epilogue.  The warpgroup that does not own the last tile takes the final reference max from the one that does
        // (consuming that tile's publication), both bring their partial row sums to it and exchange them; the owner of the last
        // tile waits for its product (it has followed that barrier), then the two meet.
        const int g_last = (T - 1) & 1;
        float m_fin = m_own;
        if (g != g_last) {
            m_fin = xm_take(xm_u32 + (uint32_t)((g_last * kFT + row) * 8), T - 1);
            const float a = ex2_approx(m_own - m_fin);           // a warpgroup without tiles (T == 1): l = 0
            l = (m_own == -INFINITY) ? 0.0f : l * a;
            mass_acc = (m_own == -INFINITY) ? 0.0f : mass_acc * a;
        } else {
            mbar_wait(pvdone_g, (uint32_t)(((T - 1) >> 1) & 1));
        }
        float *xl = reinterpret_cast<float *>(smem_k);          // all MMAs that read K have retired long before the last P V
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");          // (1) everybody is past the tile loop, the accumulator is final
        tc_fence_after();
        xl[(0 * 2 + g) * kFT + row] = l;
        xl[(1 * 2 + g) * kFT + row] = mass_acc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float L = l + xl[(0 * 2 + (g ^ 1)) * kFT + row];
        const float mass = mass_acc + xl[(1 * 2 + (g ^ 1)) * kFT + row];
        const float inv_l = 1.0f / L;
        const bool live = i < p.Sq;
        unsigned char *orow = (unsigned char *)p.out + ((size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)i * p.o_ss) * 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                            // this thread writes output columns [64g, 64g + 64)
            uint32_t o[32];
            tmem_ld32(lane_base + (uint32_t)(2 * kFT + 64 * g + 32 * c), o);
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pack2<BF16>(__uint_as_float(o[8 * q + 2 * e]) * inv_l, __uint_as_float(o[8 * q + 2 * e + 1]) * inv_l);
                    *reinterpret_cast<uint4 *>(orow + (size_t)(64 * g + 32 * c + 8 * q) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        if (live && g == 0) {
            const size_t at = ((size_t)b * p.H + h) * p.Sq + i;
            if (p.lse) p.lse[at] = (m_fin + log2f(L)) * 0.6931471805599453f;
            if (p.head_sums) p.head_sums[at] = mass * inv_l;
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- tensor-pipe rate probe (rtkv_debug_mma_rate): how many cycles does ONE 128x128x16 tcgen05.mma of the shapes used above
// take when nothing else runs?  One CTA per SM, one thread issues `n` MMAs back to back (operands: whatever is in shared /
// tensor memory), commits, waits.  mode 0: both operands in shared memory, K-major (Q K^T); 1: B MN-major (P V with P in shared
// memory); 2: A in tensor memory, B MN-major (P V with P in TMEM); 3: the real mix, 8 of mode 0 then 8 of mode 2; 4: mode 0
// with N = 256; 5: mode 0 cycling through four accumulators (no back-to-back dependency on D).
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode, int n, uint32_t idesc_qk, uint32_t idesc_pv, long long *out)
{
    extern __shared__ __align__(1024) unsigned char fp_smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)fp_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + 3 * (size_t)kFTileBytes);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);
    for (int k = threadIdx.x; k < 3 * kFTileBytes / 4; k += blockDim.x) ((uint32_t *)smem)[k] = 0x3c003c00u;   // finite values
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), bk = smem_u32(smem + kFTileBytes), bv = smem_u32(smem + 2 * (size_t)kFTileBytes);
        if (mode == 6) {                  // mode 0 with every descriptor precomputed: what does the bare instruction cost?
            uint64_t da[8], db[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t off = (uint32_t)((kk >> 2) * kFHalf + (kk & 3) * 32);
                da[kk] = umma_desc_sw128(a + off);
                db[kk] = umma_desc_sw128(bk + off);
            }
            const long long t0 = clock64();
            for (int it = 0; it < n; it += 8) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_f16(tmem_base, da[kk], db[kk], idesc_qk, 1u);
            }
            const long long t1 = clock64();
            umma_commit(smem_u32(bar));
            mbar_wait(smem_u32(bar), 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        } else {
        const long long t0 = clock64();
        for (int it = 0; it < n; ++it) {
            const int m = mode == 3 ? ((it >> 3) & 1 ? 2 : 0) : mode;
            const int kk = it & 7;
            if (m == 0) {
                const uint32_t off = (uint32_t)((kk >> 2) * kFHalf + (kk & 3) * 32);
                umma_f16(tmem_base + (uint32_t)(((it >> 4) & 1) * kFT), umma_desc_sw128(a + off), umma_desc_sw128(bk + off), idesc_qk, (uint32_t)(kk != 0));
            } else if (m == 4) {          // N = 256 (idesc_pv carries the N = 256 descriptor in this mode)
                const uint32_t off = (uint32_t)((kk & 3) * 32);
                umma_f16(tmem_base + (uint32_t)(((it >> 4) & 1) * 2 * kFT), umma_desc_sw128(a + off), umma_desc_sw128(bk + off), idesc_pv, (uint32_t)(kk != 0));
            } else if (m == 5) {          // mode 0 with the accumulator alternating every instruction
                const uint32_t off = (uint32_t)((kk >> 2) * kFHalf + (kk & 3) * 32);
                umma_f16(tmem_base + (uint32_t)((it & 3) * kFT), umma_desc_sw128(a + off), umma_desc_sw128(bk + off), idesc_qk, (uint32_t)(it >= 4));
            } else if (m == 1) {
                const uint32_t off = (uint32_t)((kk >> 2) * kFHalf + (kk & 3) * 32);
                umma_f16(tmem_base + 2 * kFT, umma_desc_sw128(a + off), umma_desc_sw128_mn(bv + (uint32_t)(kk * 16 * 128), kFHalf), idesc_pv, 1u);
            } else {
                umma_f16_ts(tmem_base + 2 * kFT, tmem_base + 3 * kFT + (uint32_t)(kk * 8), umma_desc_sw128_mn(bv + (uint32_t)(kk * 16 * 128), kFHalf), idesc_pv, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// [B,H,S,128] view with unit stride in the last dim -> 4-D tensor map {128, S, H, B}, box {64, 128, 1, 1}, 128-byte swizzle
static int encode_bhsd_map(CUtensorMap *map, const void *base, int dtype, int B, int H, int S, int64_t sb, int64_t sh, int64_t ss)
{
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) return cuda_fail(cudaErrorUnknown, "cuTensorMapEncodeTiled lookup");
    const cuuint64_t dims[4] = {128, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kFT, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, dtype == RTKV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RTKV_EINVAL; }
    return RTKV_OK;
}

} // namespace rtkv

using namespace rtkv;

static int g_flash_variant = -1;    // -1: by shape (4 from kAutoMCKeys keys on, else 0); 0: both softmax warpgroups on every tile, agreeing on the row max; 1: independent key halves; 2: two query tiles per CTA, P in TMEM; 3: alternating key tiles, two accumulators; 4: alternating key tiles, one accumulator, P in TMEM
static int g_flash_auto_keys = 3072;
static int g_flash_poly = 0;        // every g_flash_poly-th pair of exponentials of an unmasked tile on the FMA pipe (0: none)
extern "C" int rtkv_debug_flash_poly(int every)
{
    RTKV_CHECK_ARG(every == 0 || every == 2 || every == 3 || every == 4, "rtkv_debug_flash_poly: 0, 2, 3 or 4");
    g_flash_poly = every;
    return RTKV_OK;
}

extern "C" int rtkv_debug_mma_rate(int mode, int n, int ctas, void *counters, void *stream)
{
    RTKV_CHECK_ARG(mode >= 0 && mode <= 6 && n > 0 && ctas > 0 && counters, "rtkv_debug_mma_rate: bad argument");
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(rtkv::kFT >> 3) << 17) | ((uint32_t)(rtkv::kFT >> 4) << 24);
    const size_t smem = 1024 + 3 * (size_t)rtkv::kFTileBytes + 64;
    RTKV_CUDA(cudaFuncSetAttribute(rtkv::mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t idesc_n256 = (1u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | ((uint32_t)(rtkv::kFT >> 4) << 24);
    rtkv::mma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(mode, n, idesc_qk, mode == 4 ? idesc_n256 : (idesc_qk | (1u << 16)), (long long *)counters);
    RTKV_LAUNCH_CHECK("mma_rate_kernel");
    return RTKV_OK;
}

static int g_flash_skew = 0;
extern "C" int rtkv_debug_flash_skew(int cycles)
{
    RTKV_CHECK_ARG(cycles >= 0 && cycles <= 100000, "rtkv_debug_flash_skew: 0 .. 100000 cycles");
    g_flash_skew = cycles;
    return RTKV_OK;
}

static long long *g_flash_prof = nullptr;
extern "C" int rtkv_debug_flash_profile(void *counters)      // device buffer of >= 40 int64, or NULL to switch profiling off
{
    g_flash_prof = (long long *)counters;
    return RTKV_OK;
}

extern "C" int rtkv_debug_flash_variant(int variant)
{
    RTKV_CHECK_ARG(variant >= -1 && variant <= 4, "rtkv_debug_flash_variant: -1 (automatic) .. 4");
    g_flash_variant = variant;
    return RTKV_OK;
}

extern "C" int rtkv_flash_prefill(const void *Q, const void *K, const void *V, int dtype, int B, int H, int Hkv, int Sq, int Skv, int D,
                                  const int64_t *q_strides, const int64_t *k_strides, const int64_t *v_strides,
                                  float sm_scale, int causal, int P, void *out, const int64_t *o_strides,
                                  float *lse, float *head_sums, void *stream)
{
    RTKV_CHECK_ARG(Q && K && V && out && q_strides && k_strides && v_strides && o_strides, "rtkv_flash_prefill: null pointer");
    if (dtype != RTKV_BF16 && dtype != RTKV_F16) {
        set_error("rtkv_flash_prefill: Q/K/V must be bf16 or f16");
        return RTKV_EUNSUPPORTED;
    }
    if (D != 128) {
        set_error("rtkv_flash_prefill: head_dim must be 128 (got %d)", D);
        return RTKV_EUNSUPPORTED;
    }
    RTKV_CHECK_ARG(B > 0 && H > 0 && Hkv > 0 && H % Hkv == 0 && Sq > 0 && Skv > 0, "rtkv_flash_prefill: bad shape");
    RTKV_CHECK_ARG(sm_scale > 0.0f, "rtkv_flash_prefill: sm_scale must be positive");
    RTKV_CHECK_ARG(P >= 0 && P <= kFT, "rtkv_flash_prefill: prompt length must be <= 128 (got %d)", P);
    RTKV_CHECK_ARG(!head_sums || P > 0, "rtkv_flash_prefill: head_sums needs P > 0");
    for (int k = 0; k < 3; ++k) {
        RTKV_CHECK_ARG(q_strides[k] % 8 == 0 && k_strides[k] % 8 == 0 && v_strides[k] % 8 == 0 && o_strides[k] % 8 == 0,
                       "rtkv_flash_prefill: strides must be multiples of 8 elements (16 bytes)");
    }
    RTKV_CHECK_ARG((uintptr_t)Q % 16 == 0 && (uintptr_t)K % 16 == 0 && (uintptr_t)V % 16 == 0 && (uintptr_t)out % 16 == 0,
                   "rtkv_flash_prefill: tensors must be 16-byte aligned");
    CUtensorMap mq, mk, mv;
    int rc = encode_bhsd_map(&mq, Q, dtype, B, H, Sq, q_strides[0], q_strides[1], q_strides[2]);
    if (!rc) rc = encode_bhsd_map(&mk, K, dtype, B, Hkv, Skv, k_strides[0], k_strides[1], k_strides[2]);
    if (!rc) rc = encode_bhsd_map(&mv, V, dtype, B, Hkv, Skv, v_strides[0], v_strides[1], v_strides[2]);
    if (rc) return rc;
    FlashParams p;
    p.out = out; p.o_sb = o_strides[0]; p.o_sh = o_strides[1]; p.o_ss = o_strides[2];
    p.lse = lse; p.head_sums = head_sums;
    p.B = B; p.H = H; p.Hkv = Hkv; p.Sq = Sq; p.Skv = Skv; p.P = P; p.causal = causal != 0; p.is_bf16 = dtype == RTKV_BF16;
    p.scale_log2 = sm_scale * 1.4426950408889634f;
    const uint32_t fmt = dtype == RTKV_BF16 ? 1u : 0u;
    p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kFT >> 3) << 17) | ((uint32_t)(kFT >> 4) << 24);
    p.idesc_pv = p.idesc_qk | (1u << 16);                       // B operand (V) is MN-major
    p.prof = g_flash_prof;
    p.skew = g_flash_skew;
    // automatic choice: the alternating-tile kernel wins once a query tile sees enough key tiles to fill its deeper pipeline
    const int variant = g_flash_variant >= 0 ? g_flash_variant : (Skv >= g_flash_auto_keys ? 4 : 0);
    if (variant == 1) {          // independent key halves
        const int n_qt0 = (Sq + kFT - 1) / kFT;
        const long long ctas0 = (long long)n_qt0 * H * B;
        RTKV_CHECK_ARG(ctas0 < (1ll << 31), "rtkv_flash_prefill: too many tiles");
        if (dtype == RTKV_BF16) {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_ih_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIHSmem));
            flash_prefill_ih_kernel<true><<<dim3((unsigned)ctas0), kFThreads, kIHSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        } else {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_ih_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIHSmem));
            flash_prefill_ih_kernel<false><<<dim3((unsigned)ctas0), kFThreads, kIHSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        }
        RTKV_LAUNCH_CHECK("flash_prefill_ih_kernel");
        return RTKV_OK;
    }
    if (variant == 4) {          // alternating key tiles, one accumulator, P in tensor memory
        const int n_qt0 = (Sq + kFT - 1) / kFT;
        const long long ctas0 = (long long)n_qt0 * H * B;
        RTKV_CHECK_ARG(ctas0 < (1ll << 31), "rtkv_flash_prefill: too many tiles");
        void (*kmc)(const CUtensorMap, const CUtensorMap, const CUtensorMap, FlashParams) =
            dtype == RTKV_BF16 ? (g_flash_prof ? flash_prefill_mc_kernel<true, true> : flash_prefill_mc_kernel<true>) : flash_prefill_mc_kernel<false>;
        if (!g_flash_prof && g_flash_poly == 4) kmc = dtype == RTKV_BF16 ? flash_prefill_mc_kernel<true, false, 4> : flash_prefill_mc_kernel<false, false, 4>;
        if (!g_flash_prof && g_flash_poly == 3) kmc = dtype == RTKV_BF16 ? flash_prefill_mc_kernel<true, false, 3> : flash_prefill_mc_kernel<false, false, 3>;
        if (!g_flash_prof && g_flash_poly == 2) kmc = dtype == RTKV_BF16 ? flash_prefill_mc_kernel<true, false, 2> : flash_prefill_mc_kernel<false, false, 2>;
        RTKV_CUDA(cudaFuncSetAttribute(kmc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMCSmem));
        kmc<<<dim3((unsigned)ctas0), kATThreads, kMCSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        RTKV_LAUNCH_CHECK("flash_prefill_mc_kernel");
        return RTKV_OK;
    }
    if (variant == 3) {          // alternating key tiles
        const int n_qt0 = (Sq + kFT - 1) / kFT;
        const long long ctas0 = (long long)n_qt0 * H * B;
        RTKV_CHECK_ARG(ctas0 < (1ll << 31), "rtkv_flash_prefill: too many tiles");
        if (dtype == RTKV_BF16 && g_flash_prof) {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_at_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIHSmem));
            flash_prefill_at_kernel<true, true><<<dim3((unsigned)ctas0), kATThreads, kIHSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        } else if (dtype == RTKV_BF16) {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_at_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIHSmem));
            flash_prefill_at_kernel<true><<<dim3((unsigned)ctas0), kATThreads, kIHSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        } else {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_at_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIHSmem));
            flash_prefill_at_kernel<false><<<dim3((unsigned)ctas0), kATThreads, kIHSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        }
        RTKV_LAUNCH_CHECK("flash_prefill_at_kernel");
        return RTKV_OK;
    }
    if (variant == 2) {          // two query tiles per CTA, P in tensor memory
        const int n_qp = (Sq + 2 * kFT - 1) / (2 * kFT);
        const long long ctas2 = (long long)n_qp * H * B;
        RTKV_CHECK_ARG(ctas2 < (1ll << 31), "rtkv_flash_prefill: too many tiles");
        if (dtype == RTKV_BF16) {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_pp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPPSmem));
            flash_prefill_pp_kernel<true><<<dim3((unsigned)ctas2), kPPThreads, kPPSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        } else {
            RTKV_CUDA(cudaFuncSetAttribute(flash_prefill_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPPSmem));
            flash_prefill_pp_kernel<false><<<dim3((unsigned)ctas2), kPPThreads, kPPSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
        }
        RTKV_LAUNCH_CHECK("flash_prefill_pp_kernel");
        return RTKV_OK;
    }
    const int n_qt = (Sq + kFT - 1) / kFT;
    const long long ctas = (long long)n_qt * H * B;
    RTKV_CHECK_ARG(ctas < (1ll << 31), "rtkv_flash_prefill: too many tiles");
    void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, FlashParams) = nullptr;
    const bool bf = dtype == RTKV_BF16;
    switch (g_flash_poly) {
    case 2: kern = bf ? flash_prefill_kernel<true, 2> : flash_prefill_kernel<false, 2>; break;
    case 3: kern = bf ? flash_prefill_kernel<true, 3> : flash_prefill_kernel<false, 3>; break;
    case 4: kern = bf ? flash_prefill_kernel<true, 4> : flash_prefill_kernel<false, 4>; break;
    default: kern = bf ? flash_prefill_kernel<true, 0> : flash_prefill_kernel<false, 0>; break;
    }
    if (g_flash_prof && bf) kern = flash_prefill_kernel<true, 0, true>;
    RTKV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmem));
    kern<<<dim3((unsigned)ctas), kFThreads, kFSmem, (cudaStream_t)stream>>>(mq, mk, mv, p);
    RTKV_LAUNCH_CHECK("flash_prefill_kernel");
    return RTKV_OK;
}
