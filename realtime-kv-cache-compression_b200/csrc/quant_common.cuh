// quant_common.cuh — pieces shared by quant_gather.cu (reference-semantics K3) and quant_ex.cu
// (extended scale granularities / outlier split).
#pragma once
#include "common.cuh"

namespace rtkv {

constexpr int kQT = 256;     // threads per row CTA
constexpr int kMaxVec = 8;   // 16-byte vectors per thread held in registers

struct QGParams {
    const void *x[2];
    void *y[2];
    const int32_t *sel_idx;
    const rtkv_row_stats *row_stats;
    const uint8_t *labels;
    int B, S, n, out_rows, zero_pad, quantize;
    int bits[3];
    int row_begin, b_begin;   // grid offsets (rtkv_quant_gather_range)
    const float *ext_minmax;  // [B,out_rows,4] = (min0, max0, min1, max1) of the WHOLE row (head sharding) or NULL
    int neg_max;              // the max entries are stored negated (so that one MIN all-reduce combines ranks)
    int ext_by_token;         // ext_minmax is [B,S,4] indexed by TOKEN (rtkv_shard_combine) instead of by output row
    // fused peer exchange (rtkv_quant_gather_p2p): mailbox of rank s as mapped into this process
    unsigned long long mbox[8];
    unsigned *p2p_state;      // [0] epoch of the last finished call, [1] finished-CTA counter
    int rank, world, total_ctas;
};

// no dynamic indexing of the by-value parameter struct (keeps it out of local memory)
__device__ __forceinline__ int pick_bits(const QGParams &p, int label)
{
    return label == 0 ? p.bits[0] : (label == 1 ? p.bits[1] : p.bits[2]);
}
__device__ __forceinline__ const void *pick_x(const QGParams &p, int t) { return t ? p.x[1] : p.x[0]; }
__device__ __forceinline__ void *pick_y(const QGParams &p, int t) { return t ? p.y[1] : p.y[0]; }

// externally supplied whole-row min/max of (batch row b, output row r = token tok), tensor t
__device__ __forceinline__ void ext_row_minmax(const QGParams &p, int b, int r, int tok, int t, float &mn, float &mx)
{
    const size_t at = p.ext_by_token ? ((size_t)b * p.S + tok) : ((size_t)b * p.out_rows + r);
    const float *mm = p.ext_minmax + at * 4 + 2 * t;
    mn = mm[0]; mx = p.neg_max ? -mm[1] : mm[1];
}

// block-wide min / max / any-NaN over kQT threads
__device__ __forceinline__ void block_minmax(float &mn, float &mx, int nan, float *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { red[warp] = mn; red[8 + warp] = mx; }
    const int any_nan = __syncthreads_or(nan);
    float a = red[lane & 7], z = red[8 + (lane & 7)];
#pragma unroll
    for (int o = 4; o; o >>= 1) {
        a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
        z = fmaxf(z, __shfl_xor_sync(0xffffffffu, z, o));
    }
    mn = a; mx = z;
    if (any_nan) mn = mx = __int_as_float(0x7fc00000); // torch.min/max propagate NaN
}

inline int fill_params(QGParams &p, const void *X0, const void *X1, int B, int S, int n,
                       const int32_t *sel_idx, const rtkv_row_stats *row_stats, const uint8_t *labels,
                       const int32_t *bits, int out_rows, int zero_pad, void *Y0, void *Y1)
{
    p.x[0] = X0; p.x[1] = X1; p.y[0] = Y0; p.y[1] = Y1;
    p.sel_idx = sel_idx; p.row_stats = row_stats; p.labels = labels;
    p.B = B; p.S = S; p.n = n; p.out_rows = out_rows; p.zero_pad = zero_pad;
    p.quantize = (labels != nullptr && bits != nullptr);
    for (int c = 0; c < 3; ++c) p.bits[c] = bits ? bits[c] : 0;
    p.row_begin = 0; p.b_begin = 0;
    p.ext_minmax = nullptr; p.neg_max = 0; p.ext_by_token = 0;
    for (int i = 0; i < 8; ++i) p.mbox[i] = 0;
    p.p2p_state = nullptr; p.rank = 0; p.world = 1; p.total_ctas = 0;
    return RTKV_OK;
}

} // namespace rtkv

#define RTKV_DISPATCH(dt, CALL)                                 \
    switch (dt) {                                               \
    case RTKV_F32: { constexpr int DT = RTKV_F32; CALL; } break; \
    case RTKV_F16: { constexpr int DT = RTKV_F16; CALL; } break; \
    default: { constexpr int DT = RTKV_BF16; CALL; } break;      \
    }

inline int rtkv_check_bits(const int32_t *bits, int dtype, const char *who)
{
    if (!bits) return RTKV_OK;
    for (int c = 0; c < 3; ++c) {
        if (bits[c] < 1 || bits[c] > 24) { rtkv::set_error("%s: bits[%d]=%d outside 1..24", who, c, bits[c]); return RTKV_EINVAL; }
        // the reference raises here: torch.clamp(x, 0, 2^bits-1) overflows fp16 (dynamic_quantization.py:121)
        if (dtype == RTKV_F16 && bits[c] >= 16) {
            rtkv::set_error("%s: value cannot be converted to type c10::Half without overflow (bits=%d)", who, bits[c]);
            return RTKV_EUNSUPPORTED;
        }
    }
    return RTKV_OK;
}

