// capi.cu — error plumbing, device query and the host-buffer entry point of librtkv.
//
// rtkv_compress_layer_host is the whole reference call
//   RealTimePrefillCompressor.compress_layer_kv_cache (src/compression/unified_compressor.py:95-172)
// for callers that hold HOST tensors: it stages K, V and the prompt columns of
// the attention matrix to the device, runs K1a -> K2 -> K3 and copies the
// compacted result back.  K/V uploads ride a second stream so they overlap the
// scoring kernels, which only need the attention slice.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>
#include <time.h>

#include "common.cuh"

namespace rtkv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

int g_use_pdl = 1;

} // namespace rtkv

using namespace rtkv;

extern "C" int rtkv_debug_pdl(int on)
{
    g_use_pdl = on ? 1 : 0;
    return RTKV_OK;
}

extern "C" int rtkv_abi_version(void) { return RTKV_ABI_VERSION; }
extern "C" const char *rtkv_last_error(void) { return g_err; }

extern "C" int rtkv_device_info(char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("rtkv_device_info: no CUDA device (%s)", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return RTKV_ENODEVICE;
    }
    RTKV_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RTKV_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return RTKV_OK;
}

// ---- host-buffer context -----------------------------------------------------------
constexpr int kMaxPieces = 64;   // K/V upload pieces per call (per-piece events)

struct rtkv_ctx {
    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_a = nullptr, ev_sel = nullptr;
    cudaEvent_t ev_up[kMaxPieces] = {}, ev_k3[kMaxPieces] = {};
    void *dK = nullptr, *dV = nullptr, *dKo = nullptr, *dVo = nullptr, *dA = nullptr;
    size_t cap_kv = 0, cap_a = 0, cap_bs = 0;
    int cap_b = 0;
    float *mass = nullptr, *scores = nullptr;
    uint8_t *labels = nullptr, *mask = nullptr;
    int32_t *sel_idx = nullptr, *h_sel = nullptr;
    rtkv_row_stats *rs = nullptr, *h_rs = nullptr;
    rtkv_layer_stats *ls = nullptr, *h_ls = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
};

static void ctx_free_buffers(rtkv_ctx *c)
{
    void *dev[] = {c->dK, c->dV, c->dKo, c->dVo, c->dA, c->mass, c->scores, c->labels, c->mask,
                   c->sel_idx, c->rs, c->ls, c->scratch};
    for (void *p : dev) if (p) cudaFree(p);
    if (c->h_rs) cudaFreeHost(c->h_rs);
    if (c->h_ls) cudaFreeHost(c->h_ls);
    if (c->h_sel) cudaFreeHost(c->h_sel);
}

extern "C" int rtkv_ctx_create(rtkv_ctx **out)
{
    RTKV_CHECK_ARG(out, "rtkv_ctx_create: null pointer");
    int rc = rtkv_device_info(nullptr, 0, nullptr, nullptr, nullptr);
    if (rc) return rc;
    rtkv_ctx *c = new rtkv_ctx();
    cudaError_t e = cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sel, cudaEventDisableTiming);
    for (int i = 0; i < kMaxPieces && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_k3[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->ls, sizeof(rtkv_layer_stats));
    if (e == cudaSuccess) e = cudaMemset(c->ls, 0, sizeof(rtkv_layer_stats));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_ls, sizeof(rtkv_layer_stats));
    if (e != cudaSuccess) { rtkv_ctx_destroy(c); return cuda_fail(e, "rtkv_ctx_create"); }
    *out = c;
    return RTKV_OK;
}

extern "C" int rtkv_ctx_destroy(rtkv_ctx *c)
{
    if (!c) return RTKV_OK;
    cudaStream_t st[] = {c->s_main, c->s_h2d, c->s_d2h};
    for (cudaStream_t q : st) if (q) cudaStreamSynchronize(q);
    ctx_free_buffers(c);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_sel) cudaEventDestroy(c->ev_sel);
    for (int i = 0; i < kMaxPieces; ++i) {
        if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
        if (c->ev_k3[i]) cudaEventDestroy(c->ev_k3[i]);
    }
    for (cudaStream_t q : st) if (q) cudaStreamDestroy(q);
    delete c;
    return RTKV_OK;
}

template <class T> static cudaError_t regrow(T **p, size_t bytes)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
    return cudaMalloc((void **)p, bytes);
}

static int ctx_reserve(rtkv_ctx *c, size_t kv_bytes, size_t a_bytes, int B, int S)
{
    if (kv_bytes > c->cap_kv) {
        c->cap_kv = 0;
        RTKV_CUDA(regrow(&c->dK, kv_bytes)); RTKV_CUDA(regrow(&c->dV, kv_bytes));
        RTKV_CUDA(regrow(&c->dKo, kv_bytes)); RTKV_CUDA(regrow(&c->dVo, kv_bytes));
        c->cap_kv = kv_bytes;
    }
    if (a_bytes > c->cap_a) {
        c->cap_a = 0;
        RTKV_CUDA(regrow(&c->dA, a_bytes));
        c->cap_a = a_bytes;
    }
    const size_t bs = (size_t)B * S;
    if (bs > c->cap_bs) {
        c->cap_bs = 0;
        RTKV_CUDA(regrow(&c->mass, bs * 4)); RTKV_CUDA(regrow(&c->scores, bs * 4));
        RTKV_CUDA(regrow(&c->labels, bs)); RTKV_CUDA(regrow(&c->mask, bs));
        RTKV_CUDA(regrow(&c->sel_idx, bs * 4));
        if (c->h_sel) cudaFreeHost(c->h_sel);
        c->h_sel = nullptr;
        RTKV_CUDA(cudaMallocHost((void **)&c->h_sel, bs * 4));
        c->cap_bs = bs;
    }
    size_t need = 0;
    int rc = rtkv_scratch_bytes(B, S, &need);
    if (rc) return rc;
    if (need > c->scratch_bytes) {
        c->scratch_bytes = 0;
        RTKV_CUDA(regrow(&c->scratch, need));
        rc = rtkv_scratch_init(c->scratch, need, c->s_main);
        if (rc) return rc;
        c->scratch_bytes = need;
    }
    if (B > c->cap_b) {
        c->cap_b = 0;
        RTKV_CUDA(regrow(&c->rs, (size_t)B * sizeof(rtkv_row_stats)));
        if (c->h_rs) cudaFreeHost(c->h_rs);
        c->h_rs = nullptr;
        RTKV_CUDA(cudaMallocHost((void **)&c->h_rs, (size_t)B * sizeof(rtkv_row_stats)));
        c->cap_b = B;
    }
    return RTKV_OK;
}

// first output row of batch row b whose token index is >= t (sel is ascending)
static int first_row_at(const int32_t *sel, int n, int t)
{
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sel[mid] < t) lo = mid + 1; else hi = mid; }
    return lo;
}

// Three streams: uploads (attention slice first, then K/V in token-range pieces), compute,
// downloads.  Scores and the selection only need the attention slice, so they run while K/V are
// still arriving; each K/V piece is quantised and its output rows start downloading as soon as it
// lands, which keeps both PCIe directions busy at once.
extern "C" int rtkv_compress_layer_host(rtkv_ctx *c, const void *K, const void *V, int dtype, int B, int S, int n,
                                        const void *A, int a_dtype, int H, int64_t sb, int64_t sh, int64_t si,
                                        const rtkv_layer_params *prm, void *K_out, void *V_out,
                                        int32_t *max_selected, float *scores, uint8_t *labels, uint8_t *mask,
                                        rtkv_row_stats *row_stats, rtkv_layer_stats *layer_stats)
{
    RTKV_CHECK_ARG(c && K && V && A && prm && K_out && V_out && max_selected, "rtkv_compress_layer_host: null pointer");
    RTKV_CHECK_ARG(dtype_ok(dtype) && dtype_ok(a_dtype), "rtkv_compress_layer_host: bad dtype");
    RTKV_CHECK_ARG(B > 0 && S > 0 && n > 0 && H > 0, "rtkv_compress_layer_host: bad shape");
    const int P = prm->prompt_len;
    RTKV_CHECK_ARG(P > 0, "rtkv_compress_layer_host: prompt_len must be positive");
    const size_t es = dtype_bytes(dtype), ea = dtype_bytes(a_dtype);
    const size_t kv_bytes = (size_t)B * S * n * es;
    const size_t a_bytes = (size_t)B * H * S * P * ea;
    const size_t row_bytes = (size_t)n * es;
    int rc = ctx_reserve(c, kv_bytes, a_bytes, B, S);
    if (rc) return rc;
    static const bool trace = getenv("RTKV_TRACE") != nullptr;   // host-side phase times on stderr
    const auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = trace ? now() : 0.0;

    // ---- uploads: prompt columns of A (a strided 2-D copy), then K/V pieces
    if (sh == (int64_t)S * si && sb == (int64_t)H * sh) {
        RTKV_CUDA(cudaMemcpy2DAsync(c->dA, P * ea, A, (size_t)si * ea, P * ea, (size_t)B * H * S,
                                    cudaMemcpyHostToDevice, c->s_h2d));
    } else {
        for (int b = 0; b < B; ++b)
            for (int h = 0; h < H; ++h)
                RTKV_CUDA(cudaMemcpy2DAsync((char *)c->dA + ((size_t)b * H + h) * S * P * ea, P * ea,
                                            (const char *)A + ((size_t)b * sb + (size_t)h * sh) * ea,
                                            (size_t)si * ea, P * ea, (size_t)S, cudaMemcpyHostToDevice, c->s_h2d));
    }
    RTKV_CUDA(cudaEventRecord(c->ev_a, c->s_h2d));
    int per_row = kMaxPieces / B;                      // pieces per batch row
    if (per_row < 1) per_row = 1;
    int cap = 8;
    if (const char *e = getenv("RTKV_HOST_PIECES")) cap = atoi(e) > 0 ? atoi(e) : cap;   // tuning knob
    if (per_row > cap) per_row = cap;
    const int tok_per_piece = (S + per_row - 1) / per_row;
    const bool piecewise = (B * per_row <= kMaxPieces);
    int n_pieces = 0;
    if (piecewise) {
        for (int b = 0; b < B; ++b)
            for (int t0 = 0; t0 < S; t0 += tok_per_piece) {
                const int t1 = t0 + tok_per_piece < S ? t0 + tok_per_piece : S;
                const size_t off = ((size_t)b * S + t0) * row_bytes, len = (size_t)(t1 - t0) * row_bytes;
                RTKV_CUDA(cudaMemcpyAsync((char *)c->dK + off, (const char *)K + off, len, cudaMemcpyHostToDevice, c->s_h2d));
                RTKV_CUDA(cudaMemcpyAsync((char *)c->dV + off, (const char *)V + off, len, cudaMemcpyHostToDevice, c->s_h2d));
                RTKV_CUDA(cudaEventRecord(c->ev_up[n_pieces++], c->s_h2d));
            }
    } else {
        RTKV_CUDA(cudaMemcpyAsync(c->dK, K, kv_bytes, cudaMemcpyHostToDevice, c->s_h2d));
        RTKV_CUDA(cudaMemcpyAsync(c->dV, V, kv_bytes, cudaMemcpyHostToDevice, c->s_h2d));
        RTKV_CUDA(cudaEventRecord(c->ev_up[0], c->s_h2d));
    }

    // ---- scores + selection as soon as the attention slice is on the device
    RTKV_CUDA(cudaStreamWaitEvent(c->s_main, c->ev_a, 0));
    rc = rtkv_prompt_mass(c->dA, a_dtype, B, H, S, (int64_t)H * S * P, (int64_t)S * P, P, nullptr, P, c->mass, c->s_main);
    if (rc) return rc;
    rtkv_layer_params lp = *prm;
    lp.score_dtype = a_dtype;
    rc = rtkv_score_select(c->mass, c->scores, c->labels, c->mask, c->sel_idx, c->rs, c->ls, c->scratch,
                           c->scratch_bytes, B, S, &lp,
                           RTKV_FROM_MASS | RTKV_MAKE_LABELS | RTKV_SELECT | RTKV_FALLBACK, c->s_main);
    if (rc) return rc;
    RTKV_CUDA(cudaEventRecord(c->ev_sel, c->s_main));
    RTKV_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_sel, 0));
    RTKV_CUDA(cudaMemcpyAsync(c->h_ls, c->ls, sizeof(rtkv_layer_stats), cudaMemcpyDeviceToHost, c->s_d2h));
    RTKV_CUDA(cudaMemcpyAsync(c->h_rs, c->rs, (size_t)B * sizeof(rtkv_row_stats), cudaMemcpyDeviceToHost, c->s_d2h));
    if (piecewise)
        RTKV_CUDA(cudaMemcpyAsync(c->h_sel, c->sel_idx, (size_t)B * S * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    RTKV_CUDA(cudaStreamSynchronize(c->s_d2h));           // the one data-dependent shape: S' = max_selected
    const int ms = c->h_ls->max_selected;
    *max_selected = ms;
    const double t_sel = trace ? now() : 0.0;
    if (layer_stats) *layer_stats = *c->h_ls;
    if (row_stats) memcpy(row_stats, c->h_rs, (size_t)B * sizeof(rtkv_row_stats));

    // fp16 rows cannot hold 16-bit codes: the reference's torch.clamp raises (dynamic_quantization.py:121) - but only
    // for a class that HAS tokens (empty classes are skipped, :165).  An empty class's width is never used.
    if (dtype == RTKV_F16) {
        long long cnt[3] = {0, 0, 0};
        for (int b = 0; b < B; ++b)
            for (int k = 0; k < 3; ++k) cnt[k] += c->h_rs[b].label_count[k];
        for (int k = 0; k < 3; ++k) {
            if (lp.bits[k] < 16) continue;
            if (cnt[k] > 0) {
                cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_main); cudaStreamSynchronize(c->s_d2h);
                set_error("rtkv_compress_layer_host: value cannot be converted to type c10::Half without overflow (bits=%d)",
                          lp.bits[k]);
                return RTKV_EUNSUPPORTED;
            }
            lp.bits[k] = 15;
        }
    }

    // ---- quantise + download piece by piece
    if (piecewise) {
        int piece = 0;
        for (int b = 0; b < B; ++b) {
            const int cnt = c->h_rs[b].n_selected;
            const int32_t *sel = c->h_sel + (size_t)b * S;
            for (int t0 = 0; t0 < S; t0 += tok_per_piece, ++piece) {
                const int t1 = t0 + tok_per_piece < S ? t0 + tok_per_piece : S;
                const int r0 = first_row_at(sel, cnt, t0);
                const int r1 = (t1 == S) ? ms : first_row_at(sel, cnt, t1);   // the last piece carries the zero padding
                RTKV_CUDA(cudaStreamWaitEvent(c->s_main, c->ev_up[piece], 0));
                if (r1 <= r0) continue;
                rc = rtkv_quant_gather_range(c->dK, c->dV, dtype, B, S, n, c->sel_idx, c->rs, c->labels, lp.bits, ms, 1,
                                             c->dKo, c->dVo, b, 1, r0, r1 - r0, c->s_main);
                if (rc) return rc;
                RTKV_CUDA(cudaEventRecord(c->ev_k3[piece], c->s_main));
                RTKV_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k3[piece], 0));
                const size_t off = ((size_t)b * ms + r0) * row_bytes, len = (size_t)(r1 - r0) * row_bytes;
                RTKV_CUDA(cudaMemcpyAsync((char *)K_out + off, (char *)c->dKo + off, len, cudaMemcpyDeviceToHost, c->s_d2h));
                RTKV_CUDA(cudaMemcpyAsync((char *)V_out + off, (char *)c->dVo + off, len, cudaMemcpyDeviceToHost, c->s_d2h));
            }
        }
    } else {
        RTKV_CUDA(cudaStreamWaitEvent(c->s_main, c->ev_up[0], 0));
        rc = rtkv_quant_gather(c->dK, c->dV, dtype, B, S, n, c->sel_idx, c->rs, c->labels, lp.bits, ms, 1,
                               c->dKo, c->dVo, c->s_main);
        if (rc) return rc;
        RTKV_CUDA(cudaEventRecord(c->ev_k3[0], c->s_main));
        RTKV_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k3[0], 0));
        const size_t out_bytes = (size_t)B * ms * row_bytes;
        if (out_bytes) {
            RTKV_CUDA(cudaMemcpyAsync(K_out, c->dKo, out_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
            RTKV_CUDA(cudaMemcpyAsync(V_out, c->dVo, out_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
        }
    }
    const size_t bs = (size_t)B * S;
    if (scores) RTKV_CUDA(cudaMemcpyAsync(scores, c->scores, bs * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    if (labels) RTKV_CUDA(cudaMemcpyAsync(labels, c->labels, bs, cudaMemcpyDeviceToHost, c->s_d2h));
    if (mask) RTKV_CUDA(cudaMemcpyAsync(mask, c->mask, bs, cudaMemcpyDeviceToHost, c->s_d2h));
    const double t_enq = trace ? now() : 0.0;
    RTKV_CUDA(cudaStreamSynchronize(c->s_h2d));
    const double t_up = trace ? now() : 0.0;
    RTKV_CUDA(cudaStreamSynchronize(c->s_main));
    RTKV_CUDA(cudaStreamSynchronize(c->s_d2h));
    if (trace)
        fprintf(stderr, "[rtkv host] selection ready %.2f ms, enqueued %.2f, uploads done %.2f, all done %.2f (kept %d)\n",
                t_sel - t_begin, t_enq - t_begin, t_up - t_begin, now() - t_begin, ms);
    return RTKV_OK;
}

// Fused Q*K_prompt^T scoring (tcgen05): implemented in prompt_mass_qk.cu.
