// prompt_mass.cu — K1a: attention mass every query row puts on the prompt key
// columns, averaged over heads (reference: src/compression/token_importance.py:21-47,
//   prompt_attention = A[:, :, :, prompt_indices]; A_P = prompt_attention.mean(1).sum(-1)).
//
// HBM-bound streaming reduction: B*H*S*P elements are read exactly once
// (64 MiB per Llama-2-7B layer at S=4096, P=128, fp32) and B*S floats written.
//
// One CTA of 8 warps owns one query row (b,i); warp g owns head group g
// (gs = ceil(H/8) consecutive heads).  For every head of its group a warp reads
// the row's P contiguous probabilities as one 512-byte (fp32) / 256-byte (half)
// coalesced request, lane l holding columns 4l..4l+3, with all heads of the
// group in flight before the first add.  Thousands of small CTAs keep
// ~128 KB of loads in flight per SM and leave no tail wave.
//
// The reduction order is the one oracle/rtkv_oracle.c (orc_prompt_mass_groups)
// defines, so results are bit-identical to the oracle, and the fp32 per-group
// scalar partials are exactly what head-sharded ranks exchange:
//   fp32: per group: per column heads added sequentially; per lane columns in
//         increasing order; xor butterfly -> partial[g]; mass = (sum_g partial[g]) / H
//   half: per column the group sums are added in group order, m = rnd(s/H) (torch
//         rounds the head mean to the tensor dtype); lanes add m; butterfly; rnd.
#include "prompt_mass.cuh"

namespace rtkv {

// mass != NULL: full result for H heads.  partials != NULL (fp32 only): per-group
// scalar partial sums [B,S,n_groups] of the local heads (head sharding).
template <int DT, bool VEC>
__global__ void __launch_bounds__(kGroups * 32, 6)
prompt_mass_kernel(const void *__restrict__ A, int B, int H, int S, int64_t sb, int64_t sh, int64_t si,
                   const int64_t *__restrict__ idx, int P, int gs, int n_groups,
                   float *__restrict__ mass, float *__restrict__ partials)
{
    using E = typename Elem<DT>::T;
    __shared__ float part[kGroups][129];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rows = (int64_t)B * S;
    // head sharding: a rank with n_groups < 8 local groups packs 8/n_groups rows into one CTA
    const int rpc = partials ? kGroups / n_groups : 1;
    const int g = partials ? warp % n_groups : warp;
    const int sub = partials ? warp / n_groups : 0;
    const int h0 = g * gs, h1 = min(H, h0 + gs);
    const float fH = (float)H;
    pdl_wait();
    pdl_trigger();

    for (int64_t row0 = (int64_t)blockIdx.x * rpc; row0 < rows; row0 += (int64_t)gridDim.x * rpc) {
        const int64_t row = row0 + sub;
        if (row >= rows) continue;                       // (only possible in partials mode: no barrier below)
        const int b = (int)(row / S);
        const int i = (int)(row - (int64_t)b * S);
        const E *rowp = (const E *)A + b * sb + (int64_t)i * si;
        if constexpr (DT == RTKV_F32) {
            float lane_acc = 0.0f;
            for (int c0 = 0; c0 < P; c0 += 128) {
                float s[4];
                group_column_sums<DT, VEC>(rowp, sh, h0, h1, c0 + 4 * lane, P, idx, s);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + 4 * lane + k < P) lane_acc = __fadd_rn(lane_acc, s[k]);
            }
            const float pg = butterfly(lane_acc);
            if (partials) {
                if (lane == 0 && g < n_groups) partials[row * n_groups + g] = pg;
            }
            if (mass) {
                if (lane == 0) part[g][0] = pg;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float tot = part[0][0];
#pragma unroll
                    for (int k = 1; k < kGroups; ++k) tot = __fadd_rn(tot, part[k][0]);
                    mass[row] = __fdiv_rn(tot, fH);
                }
                __syncthreads();
            }
        } else {
            float lane_acc = 0.0f;
            for (int c0 = 0; c0 < P; c0 += 128) {
                float s[4];
                group_column_sums<DT, VEC>(rowp, sh, h0, h1, c0 + 4 * lane, P, idx, s);
#pragma unroll
                for (int k = 0; k < 4; ++k) part[g][4 * lane + k] = s[k];
                __syncthreads();
                if (g == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (c0 + 4 * lane + k < P) {
                            float tot = part[0][4 * lane + k];
#pragma unroll
                            for (int q = 1; q < kGroups; ++q) tot = __fadd_rn(tot, part[q][4 * lane + k]);
                            lane_acc = __fadd_rn(lane_acc, Elem<DT>::rnd(__fdiv_rn(tot, fH)));
                        }
                    }
                }
                __syncthreads();
            }
            if (g == 0) {
                const float tot = butterfly(lane_acc);
                if (lane == 0) mass[row] = Elem<DT>::rnd(tot);
            }
        }
    }
}

// combine: mass[t] = (p0 + p1 + ...) / H, groups added in head order.  partials is laid out
// [n_ranks][n][groups_per_rank] (n_ranks = 1: plain [n][n_groups]); this is exactly what
// all_gather_into_tensor of the ranks' [n][groups_per_rank] buffers produces.
__global__ void mass_from_partials_kernel(const float *__restrict__ partials, int64_t n, int n_ranks, int gpr, float fH,
                                          float *mass)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float tot = 0.0f;
    bool first = true;
    for (int r = 0; r < n_ranks; ++r) {
        const float *q = partials + ((int64_t)r * n + t) * gpr;
        for (int g = 0; g < gpr; ++g) {
            tot = first ? q[g] : __fadd_rn(tot, q[g]);
            first = false;
        }
    }
    mass[t] = __fdiv_rn(tot, fH);
}

// head-major variant for the Q.K scoring kernel: mass[b,i] = (hs[b,0,i] + hs[b,1,i] + ...) / H
__global__ void mass_from_head_sums_kernel(const float *__restrict__ hs, int H, int S, float *mass)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const float *q = hs + (size_t)b * H * S + i;
    float tot = q[0];
    for (int h = 1; h < H; ++h) tot = __fadd_rn(tot, q[(size_t)h * S]);
    mass[(size_t)b * S + i] = __fdiv_rn(tot, (float)H);
}

template <int DT>
static int launch_mass(const void *A, int B, int H, int S, int64_t sb, int64_t sh, int64_t si,
                       const int64_t *idx, int P, int gs, int n_groups, float *mass, float *partials,
                       cudaStream_t st)
{
    const int eb = Elem<DT>::kBytes;
    const int64_t quad = 4 * eb; // bytes of one lane's vector load
    const bool vec = idx == nullptr && ((uintptr_t)A % quad == 0) && (sb * eb) % quad == 0 &&
                     (sh * eb) % quad == 0 && (si * eb) % quad == 0;
    const int rpc = partials ? kGroups / n_groups : 1;
    const int64_t rows = ((int64_t)B * S + rpc - 1) / rpc;
    const int grid = (int)(rows < (1 << 20) ? rows : (1 << 20));
    if (vec)
        RTKV_CUDA(launch_kernel(prompt_mass_kernel<DT, true>, dim3(grid), dim3(kGroups * 32), 0, st, 1, A, B, H, S, sb, sh, si,
                                idx, P, gs, n_groups, mass, partials));
    else
        RTKV_CUDA(launch_kernel(prompt_mass_kernel<DT, false>, dim3(grid), dim3(kGroups * 32), 0, st, 1, A, B, H, S, sb, sh, si,
                                idx, P, gs, n_groups, mass, partials));
    return RTKV_OK;
}

} // namespace rtkv

using namespace rtkv;

extern "C" int rtkv_prompt_mass(const void *A, int dtype, int B, int H, int S,
                                int64_t sb, int64_t sh, int64_t si,
                                const int64_t *idx, int P, float *mass, void *stream)
{
    RTKV_CHECK_ARG(A && mass, "rtkv_prompt_mass: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_prompt_mass: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && H > 0 && S > 0 && P > 0, "rtkv_prompt_mass: bad shape B=%d H=%d S=%d P=%d", B, H, S, P);
    cudaStream_t st = (cudaStream_t)stream;
    const int gs = (H + kGroups - 1) / kGroups;
    switch (dtype) {
    case RTKV_F32: return launch_mass<RTKV_F32>(A, B, H, S, sb, sh, si, idx, P, gs, kGroups, mass, nullptr, st);
    case RTKV_F16: return launch_mass<RTKV_F16>(A, B, H, S, sb, sh, si, idx, P, gs, kGroups, mass, nullptr, st);
    default: return launch_mass<RTKV_BF16>(A, B, H, S, sb, sh, si, idx, P, gs, kGroups, mass, nullptr, st);
    }
}

extern "C" int rtkv_prompt_mass_partials(const void *A, int dtype, int B, int H_local, int S,
                                         int64_t sb, int64_t sh, int64_t si, const int64_t *idx, int P,
                                         int group_size, float *partials, void *stream)
{
    RTKV_CHECK_ARG(A && partials, "rtkv_prompt_mass_partials: null pointer");
    RTKV_CHECK_ARG(dtype == RTKV_F32, "rtkv_prompt_mass_partials: head sharding needs float32 probabilities "
                   "(half dtypes round the per-column head mean, which does not decompose over ranks)");
    RTKV_CHECK_ARG(B > 0 && H_local > 0 && S > 0 && P > 0 && group_size > 0, "rtkv_prompt_mass_partials: bad shape");
    const int ng = H_local / group_size;
    RTKV_CHECK_ARG(H_local % group_size == 0 && (ng == 1 || ng == 2 || ng == 4 || ng == 8),
                   "rtkv_prompt_mass_partials: H_local=%d must be 1, 2, 4 or 8 groups of group_size=%d heads",
                   H_local, group_size);
    return launch_mass<RTKV_F32>(A, B, H_local, S, sb, sh, si, idx, P, group_size, H_local / group_size, nullptr,
                                 partials, (cudaStream_t)stream);
}

extern "C" int rtkv_mass_from_partials(const float *partials, int64_t n, int n_ranks, int groups_per_rank, int H,
                                       float *mass, void *stream)
{
    RTKV_CHECK_ARG(partials && mass && n > 0 && H > 0 && n_ranks > 0 && groups_per_rank > 0,
                   "rtkv_mass_from_partials: bad argument");
    mass_from_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partials, n, n_ranks,
                                                                                          groups_per_rank, (float)H, mass);
    RTKV_LAUNCH_CHECK("mass_from_partials_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_mass_from_head_sums(const float *head_sums, int B, int H, int S, float *mass, void *stream)
{
    RTKV_CHECK_ARG(head_sums && mass && B > 0 && B <= 65535 && H > 0 && S > 0, "rtkv_mass_from_head_sums: bad argument");
    dim3 grid((S + 255) / 256, B);
    mass_from_head_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(head_sums, H, S, mass);
    RTKV_LAUNCH_CHECK("mass_from_head_sums_kernel");
    return RTKV_OK;
}
