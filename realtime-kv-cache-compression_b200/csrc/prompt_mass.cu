// prompt_mass.cu — K1a: attention mass every query row puts on the prompt key
// columns, averaged over heads (reference: src/compression/token_importance.py:21-47,
//   prompt_attention = A[:, :, :, prompt_indices]; A_P = prompt_attention.mean(1).sum(-1)).
//
// HBM-bound streaming reduction: B*H*S*P elements are read exactly once
// (64 MiB per Llama-2-7B layer at S=4096, P=128, fp32) and B*S floats written.
// One warp owns one query row (b,i).  For every head it reads the row's P
// contiguous probabilities as one 512-byte (fp32) / 256-byte (half) coalesced
// request, lane l holding columns 4l..4l+3, with eight heads' loads in flight
// before the first add.  The reduction order is the one oracle/rtkv_oracle.c
// (orc_prompt_mass) defines, so the result is bit-identical to the oracle:
//   per column: heads summed sequentially in fp32, then / H;
//   per lane  : its columns added in increasing order;
//   lanes     : xor butterfly 16,8,4,2,1.
#include "common.cuh"

namespace rtkv {

constexpr int kWarpsPerCta = 8;
constexpr int kHeadUnroll = 8;

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

// VEC: rows are 16B(f32)/8B(half)-aligned and idx == NULL -> vector loads.
template <int DT, bool VEC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
prompt_mass_kernel(const void *__restrict__ A, int B, int H, int S,
                   int64_t sb, int64_t sh, int64_t si,
                   const int64_t *__restrict__ idx, int P, float *__restrict__ mass)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t rows = (int64_t)B * S;
    const float fH = (float)H;
    using E = typename Elem<DT>::T;
    const E *base = (const E *)A;

    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp; row < rows;
         row += (int64_t)gridDim.x * kWarpsPerCta) {
        const int b = (int)(row / S);
        const int i = (int)(row - (int64_t)b * S);
        const E *rowp = base + b * sb + (int64_t)i * si;
        float lane_acc = 0.0f;
        for (int c0 = 0; c0 < P; c0 += 128) {          // chunks of 128 prompt positions
            const int p0 = c0 + 4 * lane;               // this lane's 4 positions
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (p0 < P) {
                const bool full = (p0 + 4 <= P);
                int64_t col[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int p = p0 + k;
                    col[k] = (p < P) ? (VEC ? (int64_t)p : (idx ? idx[p] : (int64_t)p)) : -1;
                }
                for (int h0 = 0; h0 < H; h0 += kHeadUnroll) {
                    float v[kHeadUnroll][4];
#pragma unroll
                    for (int u = 0; u < kHeadUnroll; ++u) {     // issue all loads first
                        const int h = h0 + u;
                        if (h < H) {
                            const E *hp = rowp + (int64_t)h * sh;
                            if (VEC && full) {
                                Quad<DT>::load_vec(hp + p0, v[u]);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    v[u][k] = col[k] >= 0 ? load_elem<DT>(hp, col[k]) : 0.0f;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[u][k] = 0.0f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kHeadUnroll; ++u)       // then add in head order
                        if (h0 + u < H) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) s[k] = __fadd_rn(s[k], v[u][k]);
                        }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (p0 + k < P)
                        lane_acc = __fadd_rn(lane_acc, Elem<DT>::rnd(__fdiv_rn(s[k], fH)));
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
            lane_acc = __fadd_rn(lane_acc, __shfl_xor_sync(0xffffffffu, lane_acc, o));
        if (lane == 0) mass[row] = Elem<DT>::rnd(lane_acc);
    }
}

template <int DT>
static int launch_mass(const void *A, int B, int H, int S, int64_t sb, int64_t sh, int64_t si,
                       const int64_t *idx, int P, float *mass, cudaStream_t st)
{
    const int eb = Elem<DT>::kBytes;
    const int64_t quad = 4 * eb; // bytes of one lane's vector load
    bool vec = idx == nullptr && ((uintptr_t)A % quad == 0) && (sb * eb) % quad == 0 &&
               (sh * eb) % quad == 0 && (si * eb) % quad == 0;
    const int64_t rows = (int64_t)B * S;
    int64_t ctas = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
    int grid = (int)(ctas < 148 * 64 ? ctas : 148 * 64);
    if (vec)
        prompt_mass_kernel<DT, true><<<grid, kWarpsPerCta * 32, 0, st>>>(A, B, H, S, sb, sh, si, idx, P, mass);
    else
        prompt_mass_kernel<DT, false><<<grid, kWarpsPerCta * 32, 0, st>>>(A, B, H, S, sb, sh, si, idx, P, mass);
    RTKV_LAUNCH_CHECK("prompt_mass_kernel");
    return RTKV_OK;
}

} // namespace rtkv

extern "C" int rtkv_prompt_mass(const void *A, int dtype, int B, int H, int S,
                                int64_t sb, int64_t sh, int64_t si,
                                const int64_t *idx, int P, float *mass, void *stream)
{
    using namespace rtkv;
    RTKV_CHECK_ARG(A && mass, "rtkv_prompt_mass: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype), "rtkv_prompt_mass: bad dtype %d", dtype);
    RTKV_CHECK_ARG(B > 0 && H > 0 && S > 0 && P > 0, "rtkv_prompt_mass: bad shape B=%d H=%d S=%d P=%d", B, H, S, P);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case RTKV_F32: return launch_mass<RTKV_F32>(A, B, H, S, sb, sh, si, idx, P, mass, st);
    case RTKV_F16: return launch_mass<RTKV_F16>(A, B, H, S, sb, sh, si, idx, P, mass, st);
    default: return launch_mass<RTKV_BF16>(A, B, H, S, sb, sh, si, idx, P, mass, st);
    }
}
