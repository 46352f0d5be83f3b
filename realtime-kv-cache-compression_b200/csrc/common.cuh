// common.cuh — shared device helpers for librtkv (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtkv.h"

namespace rtkv {

// ---- error plumbing (host) -------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define RTKV_CHECK_ARG(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) { rtkv::set_error(__VA_ARGS__); return RTKV_EINVAL; } \
    } while (0)
#define RTKV_CUDA(call)                                             \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return rtkv::cuda_fail(e__, #call); \
    } while (0)
#define RTKV_LAUNCH_CHECK(name)                                     \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return rtkv::cuda_fail(e__, name);  \
    } while (0)

// ---- element types: every arithmetic op rounds to the tensor dtype ---------
// (the reference runs torch ops on the tensor dtype; torch computes half ops in
//  fp32 and rounds the result, which is what rnd<> reproduces)
template <int DT> struct Elem;
template <> struct Elem<RTKV_F32> {
    using T = float;
    static constexpr int kBytes = 4;
    __device__ __forceinline__ static float rnd(float x) { return x; }
};
template <> struct Elem<RTKV_F16> {
    using T = __half;
    static constexpr int kBytes = 2;
    __device__ __forceinline__ static float rnd(float x) { return __half2float(__float2half_rn(x)); }
};
template <> struct Elem<RTKV_BF16> {
    using T = __nv_bfloat16;
    static constexpr int kBytes = 2;
    __device__ __forceinline__ static float rnd(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

__device__ __forceinline__ float rnd_dt(float x, int dt)
{
    if (dt == RTKV_F16) return Elem<RTKV_F16>::rnd(x);
    if (dt == RTKV_BF16) return Elem<RTKV_BF16>::rnd(x);
    return x;
}

__device__ __forceinline__ float half_bits_to_float(uint16_t b, int dt)
{
    if (dt == RTKV_BF16) return __uint_as_float((uint32_t)b << 16);
    return __half2float(__ushort_as_half(b));
}

template <int DT> __device__ __forceinline__ float load_elem(const void *p, int64_t i)
{
    if constexpr (DT == RTKV_F32) return ((const float *)p)[i];
    else if constexpr (DT == RTKV_F16) return __half2float(((const __half *)p)[i]);
    else return __bfloat162float(((const __nv_bfloat16 *)p)[i]);
}
template <int DT> __device__ __forceinline__ void store_elem(void *p, int64_t i, float v)
{
    if constexpr (DT == RTKV_F32) ((float *)p)[i] = v;
    else if constexpr (DT == RTKV_F16) ((__half *)p)[i] = __float2half_rn(v);
    else ((__nv_bfloat16 *)p)[i] = __float2bfloat16_rn(v);
}

// 128-bit streaming accesses: the K/V rows and attention slices are touched once.
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// unpack a 16-byte vector into floats (4 for f32, 8 for half types) and back
template <int DT> struct Vec16 {
    static constexpr int kElems = 16 / Elem<DT>::kBytes;
    __device__ __forceinline__ static void unpack(uint4 v, float *f)
    {
        if constexpr (DT == RTKV_F32) {
            f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
            f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
        } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if constexpr (DT == RTKV_BF16) {
                    f[2 * k] = __uint_as_float(w[k] << 16);
                    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
                } else {
                    __half2 h = *reinterpret_cast<const __half2 *>(&w[k]);
                    float2 t = __half22float2(h);
                    f[2 * k] = t.x; f[2 * k + 1] = t.y;
                }
            }
        }
    }
    __device__ __forceinline__ static uint4 pack(const float *f)
    {
        uint4 v;
        if constexpr (DT == RTKV_F32) {
            v.x = __float_as_uint(f[0]); v.y = __float_as_uint(f[1]);
            v.z = __float_as_uint(f[2]); v.w = __float_as_uint(f[3]);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if constexpr (DT == RTKV_BF16) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                    w[k] = *reinterpret_cast<uint32_t *>(&h);
                } else {
                    __half2 h = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
                    w[k] = *reinterpret_cast<uint32_t *>(&h);
                }
            }
            v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
        }
        return v;
    }
};

// ---- warp / block reductions -------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- quantisation of one element (dynamic_quantization.py:116-126) ---------------
// All operations are single IEEE roundings in the tensor dtype; the __f*_rn
// intrinsics are never contracted into FMAs.
template <int DT>
__device__ __forceinline__ float quant_code(float x, float scale, float zp, float qmax_r)
{
    float t = Elem<DT>::rnd(__fdiv_rn(x, scale));
    t = Elem<DT>::rnd(__fadd_rn(t, zp));
    float q = rintf(t);                      // round half to even, like torch.round
    // torch.clamp on the CPU is min(max(q, 0), qmax) with std::max/std::min: NaN stays NaN and -0.0 (a
    // constant row in (-0.5, 0)) stays -0.0, which fmaxf(-0.0f, 0.0f) would turn into +0.0
    q = q < 0.0f ? 0.0f : q;
    q = q > qmax_r ? qmax_r : q;
    return q;
}
template <int DT>
__device__ __forceinline__ float dequant_code(float q, float scale, float zp)
{
    return Elem<DT>::rnd(__fmul_rn(Elem<DT>::rnd(__fsub_rn(q, zp)), scale));
}
// scale / zero-point from a row's min and max (dynamic_quantization.py:79-93)
template <int DT>
__device__ __forceinline__ void quant_params(float mn, float mx, int bits, float &scale, float &zp)
{
    if (mx == mn) { scale = 1.0f; zp = 0.0f; return; }
    float qmax = (float)((1ll << bits) - 1);
    scale = Elem<DT>::rnd(__fdiv_rn(Elem<DT>::rnd(__fsub_rn(mx, mn)), qmax));
    zp = Elem<DT>::rnd(__fsub_rn(0.0f, Elem<DT>::rnd(__fdiv_rn(mn, scale))));
}
template <int DT> __device__ __forceinline__ float qmax_rounded(int bits)
{
    return Elem<DT>::rnd((float)((1ll << bits) - 1));
}

// ---- one row's quantiser with a row-level fast path ---------------------------------------------
// The quantise kernels are instruction-issue bound (ncu: K3 at ~75 % issue utilisation), and 11 of the ~25
// instructions per element are the IEEE division x/scale, whose divisor is the same for the whole row.
// Fast rows (non-constant, finite, scale and magnitudes inside [2^-30, 2^60]) use the Markstein scheme
//     y = RN(1/scale) once per row;  q0 = RN(x*y);  two FMA residual corrections
// which returns the correctly rounded quotient whenever |q0| >= 2^-30 (no underflow in the residuals;
// tools/divcheck.c: 0 mismatches in 1.9e10 guarded pairs); smaller quotients take the plain division.
// For those elements t = x/scale + zp is finite and >= +0 (x >= row min and rounding is monotone, so
// x/scale >= -zp; no intermediate overflows the tensor dtype - fp16 rows are additionally required to keep
// |x|/scale inside the half range), hence the lower clamp bound and the NaN / -0.0 cases of torch.clamp
// cannot occur and the clamp is one FMNMX.  Every other row and element runs quant_code/dequant_code unchanged.
template <int DT> struct RowQuant {
    float scale, zp, qmax_r, rcp;
    bool fast;
    __device__ __forceinline__ void init(float mn, float mx, int bits)
    {
        quant_params<DT>(mn, mx, bits, scale, zp);
        qmax_r = qmax_rounded<DT>(bits);
        fast = (mx != mn) && scale >= 0x1p-30f && scale <= 0x1p60f && fabsf(mn) <= 0x1p60f && fabsf(mx) <= 0x1p60f;
        // fp16 has a small exponent range: a narrow row far from zero makes x/scale and the zero-point overflow to inf, and
        // the reference then produces NaN (inf - inf).  Such rows must take the plain path (found by test_fake_quant_fuzz).
        if (DT == RTKV_F16) fast = fast && fabsf(mn) <= 65000.0f * scale && fabsf(mx) <= 65000.0f * scale;
        rcp = __frcp_rn(scale);
    }
    __device__ __forceinline__ float slow(float x) const
    {
        return dequant_code<DT>(quant_code<DT>(x, scale, zp, qmax_r), scale, zp);
    }
    // integer-valued code of one element (packed rows); fast rows: same guard as apply_fast
    __device__ __forceinline__ float code_slow(float x) const { return quant_code<DT>(x, scale, zp, qmax_r); }
    template <int N> __device__ __forceinline__ void codes_fast(const float (&f)[N], float (&c)[N]) const
    {
        bool ok = true;
#pragma unroll
        for (int e = 0; e < N; ++e) {
            const float q0 = __fmul_rn(f[e], rcp);
            ok = ok && (fabsf(q0) >= 0x1p-30f);
            float r = __fmaf_rn(-q0, scale, f[e]);
            const float q1 = __fmaf_rn(r, rcp, q0);
            r = __fmaf_rn(-q1, scale, f[e]);
            const float a = __fmaf_rn(r, rcp, q1);
            c[e] = fminf(rintf(Elem<DT>::rnd(__fadd_rn(Elem<DT>::rnd(a), zp))), qmax_r);
        }
        if (!ok) {
#pragma unroll
            for (int e = 0; e < N; ++e) c[e] = code_slow(f[e]);
        }
    }
    // N elements of a fast row, branch-free per element; one guard branch per vector
    template <int N> __device__ __forceinline__ void apply_fast(float (&f)[N]) const
    {
        float g[N];
        bool ok = true;
#pragma unroll
        for (int e = 0; e < N; ++e) {
            const float q0 = __fmul_rn(f[e], rcp);
            ok = ok && (fabsf(q0) >= 0x1p-30f);
            float r = __fmaf_rn(-q0, scale, f[e]);
            const float q1 = __fmaf_rn(r, rcp, q0);
            r = __fmaf_rn(-q1, scale, f[e]);
            const float a = __fmaf_rn(r, rcp, q1);                // == __fdiv_rn(x, scale) under the guard
            const float t = Elem<DT>::rnd(__fadd_rn(Elem<DT>::rnd(a), zp));
            g[e] = dequant_code<DT>(fminf(rintf(t), qmax_r), scale, zp);
        }
        if (ok) {
#pragma unroll
            for (int e = 0; e < N; ++e) f[e] = g[e];
        } else {
#pragma unroll
            for (int e = 0; e < N; ++e) f[e] = slow(f[e]);
        }
    }
};

// ---- bf16 rows: packed (two elements per instruction) fast path -------------------------------------------------
// torch computes every bf16 op in fp32 and rounds the result to bf16.  For +, -, * of two bf16 numbers that equals the
// correctly rounded bf16 operation (double rounding through a format with p' >= 2p + 2 bits is innocuous: 24 >= 2*8 + 2),
// which is what the hardware's packed HADD2 / HMUL2 .BF16 compute.  The division x/scale needs no exact fp32 quotient
// either: with 8-bit significands X, S the quotient X/S is never closer than 2^-18 (relative) to a midpoint of the bf16
// grid and never on one (2^9*X = (2j+1)*S is impossible for S < 2^8), while x * RN(1/scale) is within 2^-22 of x/scale,
// so rounding that product to bf16 gives RN_bf16(RN_f32(x/scale)) (tools/bf16_divcheck.c: exhaustive over all
// significand pairs and signs).  round-half-even of 0 <= t < 128 is (t + 128) - 128 in bf16 (ulp 1 in [128,256));
// t >= 128 is already an integer.  t >= +0 as in RowQuant, so the clamp is one packed min.
// Row guard (uniform per row): non-constant finite row, scale in [2^-30, 2^30], |min|, |max| zero or in [2^-30, 2^60]:
// every intermediate that reaches the result is then zero or a normal number, whatever the denormal mode of the packed
// units.  All other rows keep the element-wise code.  6.5 instructions per element instead of ~20.
struct Bf16RowPacked {
    float rcp;
    __nv_bfloat162 zp2, scale2, qmax2, c128;
    bool ok;
    __device__ __forceinline__ void init(float mn, float mx, float scale, float zp, float qmax_r)
    {
        const float amn = fabsf(mn), amx = fabsf(mx);
        ok = (mx != mn) && scale >= 0x1p-30f && scale <= 0x1p30f && (amn == 0.0f || (amn >= 0x1p-30f && amn <= 0x1p60f)) &&
             (amx == 0.0f || (amx >= 0x1p-30f && amx <= 0x1p60f));
        rcp = __frcp_rn(scale);
        zp2 = __float2bfloat162_rn(zp);          // zp, scale, qmax_r are bf16 values already: exact
        scale2 = __float2bfloat162_rn(scale);
        qmax2 = __float2bfloat162_rn(qmax_r);
        c128 = __float2bfloat162_rn(128.0f);
    }
    __device__ __forceinline__ uint32_t apply(uint32_t w) const
    {
        const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
        __nv_bfloat162 t = __floats2bfloat162_rn(__fmul_rn(lo, rcp), __fmul_rn(hi, rcp));     // rnd(x / scale)
        t = __hadd2(t, zp2);                                                                // rnd(. + zp)
        const __nv_bfloat162 r0 = __hsub2(__hadd2(t, c128), c128);                          // rint for t < 128
        const uint32_t big = __hge2_mask(t, c128);                                          // t >= 128: already integral
        const uint32_t tb = *reinterpret_cast<const uint32_t *>(&t), rb = *reinterpret_cast<const uint32_t *>(&r0);
        uint32_t qb = (tb & big) | (rb & ~big);
        __nv_bfloat162 q = __hmin2(*reinterpret_cast<__nv_bfloat162 *>(&qb), qmax2);        // clamp (lower bound holds)
        const __nv_bfloat162 y = __hmul2(__hsub2(q, zp2), scale2);                          // rnd(rnd(q - zp) * scale)
        return *reinterpret_cast<const uint32_t *>(&y);
    }
    __device__ __forceinline__ uint4 apply(uint4 v) const { return make_uint4(apply(v.x), apply(v.y), apply(v.z), apply(v.w)); }
};

// packed min / max of a 16-byte vector of bf16 with NaN propagation (torch.min/max propagate NaN)
__device__ __forceinline__ void bf16x8_minmax(uint4 v, __nv_bfloat162 &mn2, __nv_bfloat162 &mx2)
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[k]);
        mn2 = __hmin2_nan(mn2, h);
        mx2 = __hmax2_nan(mx2, h);
    }
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// Every hot kernel starts with pdl_wait() (the previous kernel of the stream has completed and its
// memory is visible) followed by pdl_trigger() (the next kernel may be scheduled as soon as all CTAs
// of this one are running): the successor's launch latency and ramp-up overlap this kernel's tail.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_use_pdl;   // capi.cu; rtkv_debug_pdl()

template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 int cluster_x, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_use_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster_x > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster_x; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int dtype_bytes(int dt) { return dt == RTKV_F32 ? 4 : 2; }
inline bool dtype_ok(int dt) { return dt == RTKV_F32 || dt == RTKV_F16 || dt == RTKV_BF16; }

} // namespace rtkv
