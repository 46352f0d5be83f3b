// prompt_mass.cuh — the head-group column sums shared by K1a (prompt_mass.cu) and the head-sharded
// statistics kernel (head_shard.cu).  The reduction order defined here is the one oracle/rtkv_oracle.c
// (orc_prompt_mass_groups) restates; both kernels must use it unchanged to stay bit-identical.
#pragma once
#include "common.cuh"

namespace rtkv {

constexpr int kGroups = 8;          // head groups == warps per CTA
constexpr int kHeadUnroll = 4;

template <int DT> struct Quad { // 4 consecutive elements
    __device__ __forceinline__ static void load_vec(const void *p, float f[4])
    {
        if constexpr (DT == RTKV_F32) {
            uint4 v = ldg_stream(p);
            f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
            f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
        } else {
            uint2 v = ldg_stream8(p);
            if constexpr (DT == RTKV_BF16) {
                f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
                f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
            } else {
                float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
                float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
                f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
            }
        }
    }
};

// Sum of this warp's head group for the lane's 4 columns of one 128-column chunk.
template <int DT, bool VEC>
__device__ __forceinline__ void group_column_sums(const typename Elem<DT>::T *rowp, int64_t sh, int h0, int h1,
                                                  int p0, int P, const int64_t *idx, float s[4])
{
    using E = typename Elem<DT>::T;
    s[0] = s[1] = s[2] = s[3] = 0.0f;
    if (p0 >= P) return;
    const bool full = (p0 + 4 <= P);
    int64_t col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = p0 + k;
        col[k] = (p < P) ? ((!VEC && idx) ? idx[p] : (int64_t)p) : -1;
    }
    for (int hb = h0; hb < h1; hb += kHeadUnroll) {
        float v[kHeadUnroll][4];
#pragma unroll
        for (int u = 0; u < kHeadUnroll; ++u) {             // issue every load of the batch first
            const int h = hb + u;
            if (h < h1) {
                const E *hp = rowp + (int64_t)h * sh;
                if (VEC && full) {
                    Quad<DT>::load_vec(hp + p0, v[u]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[u][k] = col[k] >= 0 ? load_elem<DT>(hp, col[k]) : 0.0f;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < kHeadUnroll; ++u)               // then add in head order
            if (hb + u < h1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] = __fadd_rn(s[k], v[u][k]);
            }
    }
}

__device__ __forceinline__ float butterfly(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

} // namespace rtkv
