// prompt_mass_qk.cu — K1b: prompt-attention mass straight from Q, K_prompt and the row
// log-sum-exp, without a materialised S x S attention matrix (the reference
// receives softmaxed attention and slices the prompt key columns, token_importance.py:41;
// the only contraction that reproduces that slice is Q . K_prompt^T).
//
//   head_sum[b,h,i] = sum_{j < P, j <= i} exp(scale * q[b,h,i,:].k[b,h,j,:] - lse[b,h,i])
//   mass[b,i]       = (head_sum[b,0,i] + ... + head_sum[b,H-1,i]) / H        (rtkv_mass_from_head_sums)
//
// This is the one dense contraction of the path, so it runs on the 5th-generation tensor
// cores: tcgen05.mma (M=128 query rows x N=128 prompt keys x K=16, bf16/f16 in, fp32
// accumulate in TMEM), operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte
// swizzle), persistent CTAs (one per SM) over work items (batch row, head, chunk of T query tiles):
//   warp 0        TMA producer: the head's K_prompt tile once per item (double buffered), Q tiles
//                 through a 4-stage ring (2 boxes of 64 columns each)
//   warp 1        TMEM allocator + MMA issuer: 8 tcgen05.mma per Q tile, tcgen05.commit -> mbarriers
//   warps 2..9    epilogue: two warpgroups, one per TMEM accumulator; tcgen05.ld 32 lanes x 32 columns
//                 software-pipelined against exp2(fma(s, c, -lse*log2e)); causal / P mask only where
//                 needed; a thread owns one query row, so row sums need no cross-warp combine
// Two 128-column accumulators in TMEM let the MMAs of tile t+1 overlap the epilogue of tile t; the
// rings keep rolling across work items, so the pipeline is filled once per CTA.
// Roofline: 2*B*H*S*P*D flop (4.3 GFLOP per 7B layer at S=4096) and B*H*S*D*2 bytes of Q
// (32 MiB): memory/MUFU bound, the tensor cores are there so the contraction is not the limiter.
#include "umma.cuh"

namespace rtkv {

constexpr int kTileM = 128;      // query rows per CTA
constexpr int kTileN = 128;      // prompt keys (P <= 128)
constexpr int kHeadDim = 128;    // K of the GEMM
constexpr int kKBlock = 64;      // elements per 128-byte swizzle atom row
constexpr int kStages = 5;       // Q-tile ring (160 KiB in flight per SM)
constexpr int kKBufs = 2;        // K_prompt tile buffers: the next work item's tile loads while this item's MMAs run
constexpr int kOperandBytes = kTileM * kHeadDim * 2;          // 32 KiB: one Q (or K) tile
constexpr int kQkThreads = 64 + 256;   // TMA warp, MMA warp, 8 epilogue warps
// smem: 2 K tiles (double buffered across work items) + kStages Q tiles + barriers + row-sum exchange
constexpr size_t kQkSmem = 1024 /*align slack*/ + (size_t)(kKBufs + kStages) * kOperandBytes + 256 + kTileM * sizeof(float);

struct QkParams {
    const float *lse;     // [B,H,S]
    float *partials;      // head sums [B,H,S]
    int B, H, S, P, gs, n_groups, causal;
    int tiles_per_item;   // query tiles that share one K_prompt tile load
    int dbg;              // bisecting aid (rtkv_debug_qk_mode): 1 no exp, 2 no TMEM loads, 3 no MMAs
    float scale_log2;     // scale * log2(e)
    uint32_t idesc;
};

__global__ void __launch_bounds__(kQkThreads, 1)
prompt_mass_qk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, QkParams p)
{
    extern __shared__ unsigned char qk_smem_raw[];
    // 128-byte swizzle needs 1024-byte aligned tiles
    unsigned char *smem = (unsigned char *)(((uintptr_t)qk_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_k = smem;                                   // [kKBufs][32 KiB]
    unsigned char *smem_q = smem + kKBufs * kOperandBytes;          // [kStages][32 KiB]
    uint64_t *bars = (uint64_t *)(smem + (size_t)(kKBufs + kStages) * kOperandBytes);
    uint64_t *qfull = bars, *qempty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = bars + 2 * kStages + 2;
    uint64_t *kfull = bars + 2 * kStages + 4, *kempty = bars + 2 * kStages + 6;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages + 8);
    float *xsum = (float *)((unsigned char *)bars + 256);        // [128 rows] row sums of the upper column half

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // persistent CTA over work items (batch row, head, chunk of T query tiles): the K_prompt tile of the
    // head is loaded once per item (double buffered), the Q tiles stream through the ring; q counts
    // Q tiles, it counts items, both across the whole life of the CTA
    const int tiles = (p.S + kTileM - 1) / kTileM;
    const int T = p.tiles_per_item;
    const int chunks = (tiles + T - 1) / T;
    const int n_items = chunks * p.H * p.B;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&qfull[s]), 1); mbar_init(smem_u32(&qempty[s]), 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull[a]), 1); mbar_init(smem_u32(&tempty[a]), 4);
            mbar_init(smem_u32(&kfull[a]), 1); mbar_init(smem_u32(&kempty[a]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // 256 TMEM columns: two 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int q = 0, it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int c = item % chunks, h = (item / chunks) % p.H, b = item / (chunks * p.H);
                const int kb = it % kKBufs;
                mbar_wait(smem_u32(&kempty[kb]), ((it / kKBufs) & 1) ^ 1);
                const uint32_t kf = smem_u32(&kfull[kb]);
                mbar_expect_tx(kf, kOperandBytes);
                const uint32_t kdst = smem_u32(smem_k + (size_t)kb * kOperandBytes);
                const int krow = (b * p.H + h) * p.P;
                tma_load_2d(kdst, &map_k, 0, krow, kf);
                tma_load_2d(kdst + kOperandBytes / 2, &map_k, kKBlock, krow, kf);
                const int t1 = min(tiles, (c + 1) * T);
                for (int tile = c * T; tile < t1; ++tile, ++q) {
                    const int s = q % kStages;
                    mbar_wait(smem_u32(&qempty[s]), ((q / kStages) & 1) ^ 1);
                    const uint32_t fb = smem_u32(&qfull[s]);
                    mbar_expect_tx(fb, kOperandBytes);
                    const uint32_t qdst = smem_u32(smem_q + (size_t)s * kOperandBytes);
                    const int qrow = (b * p.H + h) * p.S + tile * kTileM;
                    tma_load_2d(qdst, &map_q, 0, qrow, fb);
                    tma_load_2d(qdst + kOperandBytes / 2, &map_q, kKBlock, qrow, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // Descriptors are built once and advanced by one add per MMA, ring positions are counters: the issuing thread is a
            // serial resource (rebuilding both descriptors per MMA cost ~90 cycles per issue, umma.cuh / tools/mma_rate.py).
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem_q)), kd0 = umma_desc_sw128(smem_u32(smem_k));
            const uint32_t idesc = p.idesc;
            int q = 0, it = 0, s = 0;
            uint32_t sph = 0;                                      // Q ring position and phase
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int c = item % chunks;
                const int kb = it % kKBufs;
                mbar_wait(smem_u32(&kfull[kb]), (it / kKBufs) & 1);      // K_prompt tile of this head landed
                const uint64_t kd = umma_desc_advance(kd0, (uint32_t)(kb * kOperandBytes));
                const int t1 = min(tiles, (c + 1) * T);
                for (int tile = c * T; tile < t1; ++tile, ++q) {
                    const int a = q & 1;
                    mbar_wait(smem_u32(&tempty[a]), ((q >> 1) & 1) ^ 1);      // epilogue drained this accumulator
                    mbar_wait(smem_u32(&qfull[s]), sph);                     // TMA landed the Q tile
                    tc_fence_after();
                    const uint64_t qd = umma_desc_advance(qd0, (uint32_t)(s * kOperandBytes));
                    const uint32_t dcol = tmem_base + (uint32_t)(a * kTileN);
                    if (p.dbg < 3) {
#pragma unroll
                    for (int kq = 0; kq < 2; ++kq)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t off = (uint32_t)(kq * (kOperandBytes / 2) + kk * 32);   // 16 elements = 32 B
                            umma_f16(dcol, umma_desc_advance(qd, off), umma_desc_advance(kd, off), idesc, (uint32_t)((kq | kk) != 0));
                        }
                    }
                    umma_commit(smem_u32(&qempty[s]));     // Q stage reusable once these MMAs retire
                    umma_commit(smem_u32(&tfull[a]));      // accumulator ready for the epilogue
                    if (++s == kStages) { s = 0; sph ^= 1u; }
                }
                umma_commit(smem_u32(&kempty[kb]));        // every MMA that reads this K tile has retired
            }
        }
    } else {
        // ===== epilogue: two warpgroups, one per TMEM accumulator (group = (warp-2)/4 handles the Q tiles
        // with q % 2 == group); TMEM lane quadrant = warp % 4; a thread owns one query row of its tile, so
        // there is no cross-warp combine, and the two groups work on consecutive tiles concurrently =====
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int row = quad * 32 + lane;              // accumulator row == TMEM lane
        const float kLog2e = 1.4426950408889634f;
        int q = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int c = item % chunks, h = (item / chunks) % p.H, b = item / (chunks * p.H);
            const int t1 = min(tiles, (c + 1) * T);
            const float *lse_row = p.lse + ((size_t)b * p.H + h) * p.S;
            // The row's log-sum-exp is a dependent global load (~1 us under load) at the head of every tile's
            // epilogue, i.e. on the accumulator hand-off critical path (ncu: 15 % of all stall samples sat on its
            // first use).  It is therefore loaded one owned tile ahead, and only consumed after the wait for the MMA.
            float lse_next = 0.0f;
            bool have_next = false;
            for (int tile = c * T; tile < t1; ++tile, ++q) {
                if ((q & 1) != grp) continue;
                const int a = grp;
                const int i = tile * kTileM + row;
                const bool live = i < p.S;
                // rows of this tile see every prompt key (no causal / P masking needed) -> branch-free body
                const bool unmasked = (p.P == kTileN) && (!p.causal || tile * kTileM >= kTileN - 1);
                const float lse_cur = have_next ? lse_next : (live ? __ldg(lse_row + i) : 0.0f);
                have_next = tile + 2 < t1;
                if (have_next) {
                    const int in = (tile + 2) * kTileM + row;
                    lse_next = in < p.S ? __ldg(lse_row + in) : 0.0f;
                }
                mbar_wait(smem_u32(&tfull[a]), (q >> 1) & 1);
                tc_fence_after();
                const float l2 = lse_cur * kLog2e;
                const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * kTileN);
                float rs = 0.0f;
                uint32_t v0[32], v1[32];
                if (p.dbg == 0 && unmasked) {
                    // the common tile: every row sees all 128 prompt keys -> FFMA + MUFU.EX2 + FADD per element, nothing else
                    // (four partial sums break the serial FADD chain); TMEM loads one chunk ahead of the math.  ncu on the
                    // round-1 loop showed the per-element mask select executed on every tile: this path took 77.6 -> 65 us at C3.
                    float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
                    tmem_ld32_issue(tbase, v0);
#pragma unroll
                    for (int cc = 0; cc < kTileN; cc += 64) {
                        tmem_ld_wait();
                        tmem_ld32_issue(tbase + cc + 32, v1);
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            r0 += ex2_approx(__fmaf_rn(__uint_as_float(v0[e]), p.scale_log2, -l2));
                            r1 += ex2_approx(__fmaf_rn(__uint_as_float(v0[e + 1]), p.scale_log2, -l2));
                            r2 += ex2_approx(__fmaf_rn(__uint_as_float(v0[e + 2]), p.scale_log2, -l2));
                            r3 += ex2_approx(__fmaf_rn(__uint_as_float(v0[e + 3]), p.scale_log2, -l2));
                        }
                        tmem_ld_wait();
                        if (cc + 64 < kTileN) tmem_ld32_issue(tbase + cc + 64, v0);
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            r0 += ex2_approx(__fmaf_rn(__uint_as_float(v1[e]), p.scale_log2, -l2));
                            r1 += ex2_approx(__fmaf_rn(__uint_as_float(v1[e + 1]), p.scale_log2, -l2));
                            r2 += ex2_approx(__fmaf_rn(__uint_as_float(v1[e + 2]), p.scale_log2, -l2));
                            r3 += ex2_approx(__fmaf_rn(__uint_as_float(v1[e + 3]), p.scale_log2, -l2));
                        }
                    }
                    rs = (r0 + r1) + (r2 + r3);
                } else if (p.dbg < 2 || p.dbg == 3) {
                    // masked tiles (the first query tile under a causal mask, P < 128) and the bisecting modes
                    tmem_ld32_issue(tbase, v0);
#pragma unroll
                    for (int cc = 0; cc < kTileN; cc += 64) {
                        tmem_ld_wait();
                        tmem_ld32_issue(tbase + cc + 32, v1);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int j = cc + e;
                            float x = __uint_as_float(v0[e]);
                            if (p.dbg != 1) x = ex2_approx(__fmaf_rn(x, p.scale_log2, -l2));
                            rs += (unmasked || ((j < p.P) && (!p.causal || j <= i))) ? x : 0.0f;
                        }
                        tmem_ld_wait();
                        if (cc + 64 < kTileN) tmem_ld32_issue(tbase + cc + 64, v0);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int j = cc + 32 + e;
                            float x = __uint_as_float(v1[e]);
                            if (p.dbg != 1) x = ex2_approx(__fmaf_rn(x, p.scale_log2, -l2));
                            rs += (unmasked || ((j < p.P) && (!p.causal || j <= i))) ? x : 0.0f;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tempty[a]));
                if (live) p.partials[((size_t)b * p.H + h) * p.S + i] = rs;       // head-major: coalesced
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

static int encode_rows_map(CUtensorMap *map, const void *base, int dtype, uint64_t rows)
{
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) return cuda_fail(cudaErrorUnknown, "cuTensorMapEncodeTiled lookup");
    const cuuint64_t dims[2] = {(cuuint64_t)kHeadDim, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kHeadDim * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kKBlock, (cuuint32_t)kTileM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == RTKV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RTKV_EINVAL; }
    return RTKV_OK;
}

} // namespace rtkv

using namespace rtkv;

static int g_qk_dbg = 0;
extern "C" int rtkv_debug_qk_mode(int mode) { g_qk_dbg = mode; return RTKV_OK; }

extern "C" int rtkv_prompt_mass_qk(const void *Q, const void *Kp, const float *lse, int dtype, int B, int H, int S,
                                   int D, int P, float scale, int causal, float *head_sums, void *stream)
{
    RTKV_CHECK_ARG(Q && Kp && lse && head_sums, "rtkv_prompt_mass_qk: null pointer");
    if (dtype != RTKV_BF16 && dtype != RTKV_F16) {
        set_error("rtkv_prompt_mass_qk: Q/K must be bf16 or f16 (tf32 would cost 13 mantissa bits of the logits)");
        return RTKV_EUNSUPPORTED;
    }
    if (D != kHeadDim || P < 1 || P > kTileN) {
        set_error("rtkv_prompt_mass_qk: needs head_dim == 128 and 1 <= P <= 128 (got D=%d, P=%d)", D, P);
        return RTKV_EUNSUPPORTED;
    }
    RTKV_CHECK_ARG(B > 0 && H > 0 && S > 0, "rtkv_prompt_mass_qk: bad shape");
    RTKV_CHECK_ARG((int64_t)B * H * S < (1ll << 31), "rtkv_prompt_mass_qk: B*H*S too large");
    RTKV_CHECK_ARG((uintptr_t)Q % 16 == 0 && (uintptr_t)Kp % 16 == 0, "rtkv_prompt_mass_qk: Q/K must be 16-byte aligned");
    CUtensorMap map_q, map_k;
    int rc = encode_rows_map(&map_q, Q, dtype, (uint64_t)B * H * S);
    if (rc) return rc;
    rc = encode_rows_map(&map_k, Kp, dtype, (uint64_t)B * H * P);
    if (rc) return rc;
    static int sm_counts[64] = {0};                 // per device (a process may drive several GPUs)
    int dev = 0;
    RTKV_CUDA(cudaGetDevice(&dev));
    int sm_count = (dev >= 0 && dev < 64) ? sm_counts[dev] : 0;
    if (!sm_count) {
        RTKV_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_counts[dev] = sm_count;
    }
    QkParams p;
    p.lse = lse; p.partials = head_sums;
    p.B = B; p.H = H; p.S = S; p.P = P; p.gs = 1; p.n_groups = H; p.causal = causal;
    p.scale_log2 = scale * 1.4426950408889634f;
    const uint32_t fmt = dtype == RTKV_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    // Query tiles per work item: as many as possible (each item loads its head's K_prompt tile once)
    // while keeping at least ~1.5 items per SM for balance.
    // Query tiles per work item (each item loads its head's K_prompt tile once): the power of two that minimises the busiest
    // SM's tile count, an item switch charged as half a tile (round 1 required >= 1.5 items per SM, which left 256 items of 32
    // tiles on 148 SMs at 32k tokens: the busiest SM did 64 tiles against an average of 55).
    const int tiles = (S + kTileM - 1) / kTileM;
    int T = 1;
    double best = 1e30;
    for (int t = 1; t <= tiles; t *= 2) {
        const long long items = (long long)((tiles + t - 1) / t) * H * B;
        const long long per_sm = (items + sm_count - 1) / sm_count;
        const double cost = (double)per_sm * (t + 0.5);
        if (cost < best) { best = cost; T = t; }
    }
    p.tiles_per_item = T;
    p.dbg = g_qk_dbg;
    const long long n_items = (long long)((tiles + T - 1) / T) * H * B;
    RTKV_CHECK_ARG(n_items < (1ll << 31), "rtkv_prompt_mass_qk: too many work items");
    RTKV_CUDA(cudaFuncSetAttribute(prompt_mass_qk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQkSmem));
    dim3 grid((unsigned)(n_items < sm_count ? n_items : sm_count));      // persistent: one CTA per SM
    prompt_mass_qk_kernel<<<grid, kQkThreads, kQkSmem, (cudaStream_t)stream>>>(map_q, map_k, p);
    RTKV_LAUNCH_CHECK("prompt_mass_qk_kernel");
    return RTKV_OK;
}
