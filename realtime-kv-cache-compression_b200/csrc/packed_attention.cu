// packed_attention.cu — decode-step attention directly over packed K/V rows.
//
// The reference hands the compressed K/V back as the layer's cache entry (src/models/modified_llama.py:165-166) as
// full-width floats; with rtkv_quant_pack the kept rows live as 8/4/2-bit codes + one (scale, zero-point) per row,
// and this kernel is the consumer that never expands them in memory:
//     out[b,h,:] = sum_j softmax_j( sm_scale * q[b,h,:].K'[b,j,hk,:] ) * V'[b,j,hk,:],   K' = (code - zp) * scale
// Decode attention is bandwidth bound on K'/V', so reading codes instead of floats is the whole point: HBM bytes per
// kept token and kv head are 2*D*bits/8 (+16 for the parameters) instead of 2*D*4.
//
// Layout of the work: grid (H, splits, B) - heads fastest, so the CTAs running together read neighbouring pieces of the
// same packed rows.  A warp takes 32 kept rows per step:
//   scores  each LANE owns one row: it loads that row's 128*bits/8 bytes of this head with 16-byte loads and does the
//           whole dot product in-lane against q in shared memory (no per-row shuffles), folded dequantisation
//               q.K'_j = scale_j * (sum_d q_d*code_jd - zp_j * sum_d q_d)
//   softmax one warp max and one warp sum per 32 rows (online, log2 domain)
//   values  channel-per-lane: for each of the 32 rows p_j*scale_j is broadcast by shuffle, a lane loads its 4 codes of the
//           row (a warp reads the row's 128*bits/8 contiguous bytes) and does 4 FMAs;
//               sum_j p_j V'_j = sum_j (p_j*scale_j) * code_j - sum_j p_j*scale_j*zp_j
// The warps' states are merged through shared memory, the split's partial result goes to a workspace that
// packed_attention_combine_kernel merges (flash-decoding).  First version (warp per row for both phases): 322 us at 32k
// rows x 32 heads; this one: see DESIGN.md.
// Project-defined extension: checked against attention over PackedKV.dequantize() (tests/test_packed_attention_gpu.py).
#include "common.cuh"

namespace rtkv {

constexpr int kPaThreads = 128;
constexpr int kPaWarps = kPaThreads / 32;
constexpr int kPaD = 128;
// per-warp staging of one 32-row step (sized for 8-bit rows): K slices at a padded stride (a lane reads its own row with
// 16-byte loads: stride 144 keeps the 8 lanes of a wavefront on disjoint banks), V slices back to back
constexpr int kPaKStride = 144, kPaVStride = 128;
constexpr int kPaStageBytes = 32 * kPaKStride + 32 * kPaVStride;          // 8704
constexpr int kPaSmemBytes = kPaWarps * kPaStageBytes;                     // one K and one V buffer per warp: 34816

struct PaParams {
    const float *q;            // [B,H,D]
    const uint8_t *ck, *cv;    // [B,out_rows,row_bytes]
    const float *pk, *pv;      // [B,out_rows,2] (scale, zero-point)
    const uint8_t *row_bits;   // [B,out_rows] bit width of each packed row
    const rtkv_row_stats *row_stats;
    int64_t row_bytes;
    int B, H, Hkv, out_rows, splits;
    float sm_scale_log2;       // softmax scale * log2(e)
    float *ws;                 // [B,H,splits,2+D]  (m, l, acc[D]) in the log2 domain
    float *out;                // [B,H,D]
    float *lse;                // [B,H] natural log, optional
};

__device__ __forceinline__ unsigned load_codes4(const uint8_t *row, int c0, int bits)
{
    const uint8_t *p = row + ((size_t)c0 * bits >> 3);
    if (bits == 8) return *(const unsigned *)p;
    if (bits == 4) return *(const uint16_t *)p;
    if (bits == 2) return *p;
    return (unsigned)(*p >> (c0 & 7));          // 1 bit: 4 codes in a nibble
}

// Exact int -> float without the quarter-rate I2F pipe and without shifts: for a code sitting at bit position pos < 16 of
// x, (x & (MASK << pos)) | 0x4b000000 is the float 2^23 + code*2^pos, so one LOP3 and one FADD give code*2^pos exactly.
// The factor 2^pos is compensated by the other operand: q pre-scaled by 2^-pos in shared memory for the scores, one
// accumulator set per bit width rescaled once at the end for the values.  (ncu on the first lane-per-row version:
// ALU pipe 67 %, FMA pipe 25 % - shift+mask per code and run-time bit widths were the limiter.)
__device__ __forceinline__ float code_scaled(unsigned x, unsigned mask_at_pos)
{
    unsigned y;                                   // (x & mask) | 0x4b000000 as ONE 3-input LOP3 (both constants in registers)
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(y) : "r"(x), "r"(mask_at_pos), "r"(0x4b000000u));
    return __uint_as_float(y) - 8388608.0f;
}
__device__ __forceinline__ float code_f(unsigned word, int shift, unsigned mask)      // generic (mixed bit widths)
{
    return __uint_as_float(((word >> shift) & mask) | 0x4b000000u) - 8388608.0f;
}

// sum_d q[d] * code[d] over the 128 channels of one head of one packed row, read from the staged copy in shared memory
// (BITS x 16-byte loads).  qs = q pre-scaled for this bit width: qs[d] = q[d] * 2^-((d % (16/BITS)) * BITS)
template <int BITS>
__device__ __forceinline__ float dot_row(const uint8_t *kp, const float *qs)
{
    constexpr int PER_HALF = 16 / BITS;             // codes per 16-bit half word
    constexpr unsigned MASK = (1u << BITS) - 1u;
    float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int c = 0; c < BITS; ++c) {
        const uint4 w = *((const uint4 *)kp + c);
        const unsigned word[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const unsigned x = hh ? word[k] >> 16 : word[k];
                const float *qd = qs + ((c * 4 + k) * 2 + hh) * PER_HALF;
#pragma unroll
                for (int i = 0; i < PER_HALF; ++i) d[i & 3] = fmaf(qd[i], code_scaled(x, MASK << (i * BITS)), d[i & 3]);
            }
        }
    }
    return (d[0] + d[1]) + (d[2] + d[3]);
}

// values of `nrows` staged rows that all use BITS-bit codes: acc[e] accumulates w_j * code * 2^pos(e).
// c4 = first of this lane's 4 channels inside the head (4 * lane)
template <int BITS>
__device__ __forceinline__ void value_rows(const uint8_t *sv, int c4, int nrows, const float *sw, float (&acc)[4])
{
    constexpr unsigned MASK = (1u << BITS) - 1u;
#pragma unroll 8
    for (int j = 0; j < nrows; ++j) {
        const unsigned wv = load_codes4(sv + j * kPaVStride, c4, BITS);
        const float w = sw[j];
        if (BITS == 8) {
            const unsigned hi = wv >> 16;
            acc[0] = fmaf(w, code_scaled(wv, 0xffu), acc[0]);
            acc[1] = fmaf(w, code_scaled(wv, 0xff00u), acc[1]);
            acc[2] = fmaf(w, code_scaled(hi, 0xffu), acc[2]);
            acc[3] = fmaf(w, code_scaled(hi, 0xff00u), acc[3]);
        } else {
            acc[0] = fmaf(w, code_scaled(wv, MASK), acc[0]);
            acc[1] = fmaf(w, code_scaled(wv, MASK << BITS), acc[1]);
            acc[2] = fmaf(w, code_scaled(wv, MASK << (2 * BITS)), acc[2]);
            acc[3] = fmaf(w, code_scaled(wv, MASK << (3 * BITS)), acc[3]);
        }
    }
}

__device__ __forceinline__ void pa_cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(kPaThreads) packed_attention_split_kernel(PaParams p)
{
    extern __shared__ __align__(16) unsigned char pa_smem[];
    __shared__ __align__(16) float s_q[4][kPaD];              // q pre-scaled for 1/2/4/8-bit rows
    __shared__ float s_m[kPaWarps], s_l[kPaWarps], s_acc[kPaWarps][kPaD];
    __shared__ __align__(16) float s_w[kPaWarps][32];
    const int h = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    const int cnt = min(p.row_stats[b].n_selected, p.out_rows);
    const int per = (((cnt + p.splits - 1) / p.splits) + kPaThreads - 1) / kPaThreads * kPaThreads;   // equal steps per warp
    const int r0 = split * per, r1 = min(cnt, r0 + per);
    const int hk = h / (p.H / p.Hkv);
    float qsum = 0.0f;
    {
        const int d = threadIdx.x;                          // kPaThreads == kPaD
        const float qd = p.q[((size_t)b * p.H + h) * kPaD + d];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int bits = 1 << t;
            s_q[t][d] = qd * exp2f(-(float)((d % (16 / bits)) * bits));
        }
    }
    __syncthreads();
    {
        float part = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int d = 4 * lane + e; part += s_q[3][d] * ((d & 1) ? 256.0f : 1.0f); }
        qsum = part;
#pragma unroll
        for (int o = 16; o; o >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
    }

    float m = -INFINITY, l = 0.0f, zacc = 0.0f;
    // one accumulator set per bit width (implicit scale 2^pos) + one for steps with mixed widths
    float acc1[4] = {0, 0, 0, 0}, acc2[4] = {0, 0, 0, 0}, acc4[4] = {0, 0, 0, 0}, acc8[4] = {0, 0, 0, 0}, accg[4] = {0, 0, 0, 0};
    const size_t base = (size_t)b * p.out_rows;
    unsigned char *sk_w = pa_smem + (size_t)warp * kPaStageBytes;            // this warp's K buffer (32 rows)
    unsigned char *sv_w = sk_w + 32 * kPaKStride;                           // and V buffer
    const int step_stride = kPaWarps * 32;

    // Software pipeline per warp with ONE K and ONE V buffer: the K slices of step n+1 are copied global -> shared with
    // cp.async as soon as the dot products of step n have consumed the buffer, the V slices as soon as the value phase of
    // step n is done; each copy overlaps the other half of the step (and the other warps).  A lane copies its own row's
    // bits*16 bytes; the per-row metadata runs ahead because the copy addresses depend on the bit width.
    struct Meta { int bits; float2 kp, vp; };
    auto load_meta = [&](int rb) {
        Meta mt; mt.bits = 0; mt.kp = make_float2(0.0f, 0.0f); mt.vp = mt.kp;
        if (rb < r1) {
            const size_t row = base + (rb + lane < r1 ? rb + lane : rb);
            mt.bits = p.row_bits[row];
            mt.kp = *(const float2 *)(p.pk + row * 2);
            mt.vp = *(const float2 *)(p.pv + row * 2);
        }
        return mt;
    };
    auto issue_copy = [&](const uint8_t *codes, unsigned char *buf, int stride, int rb, int bits) {
        if (rb + lane < r1) {
            const size_t off = (base + rb + lane) * (size_t)p.row_bytes + (((size_t)hk * kPaD * bits) >> 3);
            for (int c = 0; c < bits; ++c) pa_cp_async16(buf + lane * stride + 16 * c, codes + off + 16 * c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");    // empty group past the end: keeps the accounting uniform
    };
    int rb = r0 + warp * 32;
    Meta cur = load_meta(rb), nxt = load_meta(rb + step_stride);
    issue_copy(p.ck, sk_w, kPaKStride, rb, cur.bits);           // commit order from here on: K(n), V(n), K(n+1), V(n+1), ...
    issue_copy(p.cv, sv_w, kPaVStride, rb, cur.bits);
    for (; rb < r1; rb += step_stride) {
        const bool live = rb + lane < r1;
        const int nrows = min(32, r1 - rb);
        const int bits = cur.bits;
        const float2 kp = cur.kp, vp = cur.vp;
        asm volatile("cp.async.wait_group 1;" ::: "memory");    // K(n) landed (V(n) may still fly); a lane reads its own row only
        const unsigned char *sk = sk_w + lane * kPaKStride;
        const unsigned char *sv = sv_w;
        // bit width of every row of the step from two ballots (log2(bits) in 0..3)
        const int lg = 31 - __clz(bits);
        const unsigned b0 = __ballot_sync(0xffffffffu, lg & 1), b1 = __ballot_sync(0xffffffffu, lg & 2);
        // ---- scores: lane = row
        float dot = 0.0f;
        if (live) {
            if (bits == 4) dot = dot_row<4>(sk, s_q[2]);
            else if (bits == 8) dot = dot_row<8>(sk, s_q[3]);
            else if (bits == 2) dot = dot_row<2>(sk, s_q[1]);
            else dot = dot_row<1>(sk, s_q[0]);
        }
        const float s = live ? p.sm_scale_log2 * kp.x * (dot - kp.y * qsum) : -INFINITY;
        issue_copy(p.ck, sk_w, kPaKStride, rb + step_stride, nxt.bits);      // K(n+1): the K buffer is free again
        // ---- online softmax over the 32 rows of the step
        float smax = s;
#pragma unroll
        for (int o = 16; o; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        const float m_new = fmaxf(m, smax);
        const float corr = exp2f(m - m_new);
        const float pj = live ? exp2f(s - m_new) : 0.0f;
        const float wj = pj * vp.x;                            // p_j * scale_j of this lane's row
        s_w[warp][lane] = wj;
        float psum = pj, zsum = wj * vp.y;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            psum += __shfl_xor_sync(0xffffffffu, psum, o);
            zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
        }
        l = fmaf(l, corr, psum);
        zacc = fmaf(zacc, corr, zsum);
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc1[e] *= corr; acc2[e] *= corr; acc4[e] *= corr; acc8[e] *= corr; accg[e] *= corr; }
        m = m_new;
        asm volatile("cp.async.wait_group 1;" ::: "memory");    // V(n) landed (K(n+1) may still fly)
        __syncwarp();                                           // every lane's V row and weight are visible to the warp
        // ---- values: lane = 4 channels; p_j * scale_j from shared memory (broadcast reads)
        const bool uniform = (b0 == 0u || b0 == 0xffffffffu) && (b1 == 0u || b1 == 0xffffffffu);
        if (uniform) {
            if (bits == 4) value_rows<4>(sv, 4 * lane, nrows, s_w[warp], acc4);
            else if (bits == 8) value_rows<8>(sv, 4 * lane, nrows, s_w[warp], acc8);
            else if (bits == 2) value_rows<2>(sv, 4 * lane, nrows, s_w[warp], acc2);
            else value_rows<1>(sv, 4 * lane, nrows, s_w[warp], acc1);
        } else {
            for (int j = 0; j < nrows; ++j) {
                const int bj = 1 << ((((b1 >> j) & 1) << 1) | ((b0 >> j) & 1));
                const unsigned mask = (1u << bj) - 1u;
                const unsigned wv = load_codes4(sv + j * kPaVStride, 4 * lane, bj);
                const float w = s_w[warp][j];
                accg[0] = fmaf(w, code_f(wv, 0, mask), accg[0]);
                accg[1] = fmaf(w, code_f(wv, bj, mask), accg[1]);
                accg[2] = fmaf(w, code_f(wv, 2 * bj, mask), accg[2]);
                accg[3] = fmaf(w, code_f(wv, 3 * bj, mask), accg[3]);
            }
        }
        __syncwarp();                                           // all lanes are done with the V buffer and the weights
        issue_copy(p.cv, sv_w, kPaVStride, rb + step_stride, nxt.bits);      // V(n+1)
        cur = nxt;
        nxt = load_meta(rb + 2 * step_stride);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // undo the implicit 2^pos of the per-width accumulators
    float acc[4];
    acc[0] = accg[0] + acc8[0] + acc4[0] + acc2[0] + acc1[0];
    acc[1] = accg[1] + acc8[1] * 0x1p-8f + acc4[1] * 0x1p-4f + acc2[1] * 0x1p-2f + acc1[1] * 0x1p-1f;
    acc[2] = accg[2] + acc8[2] + acc4[2] * 0x1p-8f + acc2[2] * 0x1p-4f + acc1[2] * 0x1p-2f;
    acc[3] = accg[3] + acc8[3] * 0x1p-8f + acc4[3] * 0x1p-12f + acc2[3] * 0x1p-6f + acc1[3] * 0x1p-3f;
    // merge the warps of the CTA
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
    for (int e = 0; e < 4; ++e) s_acc[warp][4 * lane + e] = acc[e] - zacc;
    __syncthreads();
    if (threadIdx.x < kPaD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kPaWarps; ++w) M = fmaxf(M, s_m[w]);
        float L = 0.0f, A = 0.0f;
#pragma unroll
        for (int w = 0; w < kPaWarps; ++w) {
            const float c = s_m[w] == -INFINITY ? 0.0f : exp2f(s_m[w] - M);
            L += s_l[w] * c; A += s_acc[w][threadIdx.x] * c;
        }
        float *o = p.ws + (((size_t)b * p.H + h) * p.splits + split) * (2 + kPaD);
        if (threadIdx.x == 0) { o[0] = M; o[1] = L; }
        o[2 + threadIdx.x] = A;
    }
}

__global__ void __launch_bounds__(kPaD) packed_attention_combine_kernel(PaParams p)
{
    const int h = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    pdl_trigger();
    const float *w = p.ws + ((size_t)b * p.H + h) * p.splits * (2 + kPaD);
    float M = -INFINITY;
    for (int s = 0; s < p.splits; ++s) M = fmaxf(M, w[(size_t)s * (2 + kPaD)]);
    float L = 0.0f, A = 0.0f;
    for (int s = 0; s < p.splits; ++s) {
        const float ms = w[(size_t)s * (2 + kPaD)];
        const float c = ms == -INFINITY ? 0.0f : exp2f(ms - M);
        L += w[(size_t)s * (2 + kPaD) + 1] * c;
        A += w[(size_t)s * (2 + kPaD) + 2 + threadIdx.x] * c;
    }
    p.out[((size_t)b * p.H + h) * kPaD + threadIdx.x] = L > 0.0f ? A / L : 0.0f;
    if (p.lse && threadIdx.x == 0) p.lse[(size_t)b * p.H + h] = L > 0.0f ? (M + log2f(L)) * 0.6931471805599453f : -INFINITY;
}

} // namespace rtkv

using namespace rtkv;

extern "C" int64_t rtkv_packed_attention_workspace_bytes(int B, int H, int splits)
{
    if (B <= 0 || H <= 0 || splits <= 0) return -1;
    return (int64_t)B * H * splits * (2 + kPaD) * (int64_t)sizeof(float);
}

extern "C" int rtkv_packed_decode_attention(const float *q, const uint8_t *codes_k, const float *qparams_k,
                                            const uint8_t *codes_v, const float *qparams_v, const uint8_t *row_bits,
                                            const rtkv_row_stats *row_stats, int64_t row_bytes, int B, int H, int Hkv,
                                            int D, int out_rows, float sm_scale, int splits, void *workspace,
                                            size_t workspace_bytes, float *out, float *lse, void *stream)
{
    RTKV_CHECK_ARG(q && codes_k && qparams_k && codes_v && qparams_v && row_bits && row_stats && workspace && out,
                   "rtkv_packed_decode_attention: null pointer");
    RTKV_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && Hkv > 0 && H % Hkv == 0 && out_rows > 0,
                   "rtkv_packed_decode_attention: bad shape");
    if (D != kPaD) { set_error("rtkv_packed_decode_attention: head dimension %d (only 128 is implemented)", D); return RTKV_EUNSUPPORTED; }
    RTKV_CHECK_ARG(splits > 0 && splits <= 65535, "rtkv_packed_decode_attention: splits outside 1..65535");
    RTKV_CHECK_ARG(row_bytes % 16 == 0 && row_bytes >= (int64_t)Hkv * D / 8, "rtkv_packed_decode_attention: bad row_bytes");
    RTKV_CHECK_ARG((int64_t)workspace_bytes >= rtkv_packed_attention_workspace_bytes(B, H, splits),
                   "rtkv_packed_decode_attention: workspace too small");
    RTKV_CHECK_ARG((uintptr_t)q % 16 == 0 && (uintptr_t)codes_k % 16 == 0 && (uintptr_t)codes_v % 16 == 0 &&
                   (uintptr_t)qparams_k % 8 == 0 && (uintptr_t)qparams_v % 8 == 0, "rtkv_packed_decode_attention: misaligned pointer");
    PaParams p = {};
    p.q = q; p.ck = codes_k; p.cv = codes_v; p.pk = qparams_k; p.pv = qparams_v; p.row_bits = row_bits;
    p.row_stats = row_stats; p.row_bytes = row_bytes; p.B = B; p.H = H; p.Hkv = Hkv; p.out_rows = out_rows;
    p.splits = splits; p.sm_scale_log2 = sm_scale * 1.4426950408889634f; p.ws = (float *)workspace; p.out = out; p.lse = lse;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_done = false;
    if (!attr_done) {
        RTKV_CUDA(cudaFuncSetAttribute(packed_attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPaSmemBytes));
        attr_done = true;
    }
    RTKV_CUDA(launch_kernel(packed_attention_split_kernel, dim3(H, splits, B), dim3(kPaThreads), kPaSmemBytes, st, 1, p));
    RTKV_CUDA(launch_kernel(packed_attention_combine_kernel, dim3(H, B), dim3(kPaD), 0, st, 1, p));
    RTKV_LAUNCH_CHECK("rtkv_packed_decode_attention");
    return RTKV_OK;
}
