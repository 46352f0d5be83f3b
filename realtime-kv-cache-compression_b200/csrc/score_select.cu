// score_select.cu — K2: importance scores -> precision labels -> budgeted token
// selection -> compacted index list + statistics, one CTA per batch row.
//
// Reference semantics (paths relative to the reference repo):
//   scores  src/compression/token_importance.py:49-176   s = ((a*n)*w_l + b*pb) + g*r
//   labels  src/compression/dynamic_quantization.py:38-45
//   select  src/compression/selective_propagation.py:40-161  greedy walk over the
//           descending score order, token kept iff cur + cost <= budget, never stops
//   fallback selective_propagation.py:203-211 (top max(1,int(0.1*S)) when nothing kept)
//
// The reference walks a sorted list one token at a time.  Here the walk is
// restated so that no sort of the row is needed:
//   * cur only grows, so once a token of cost c is rejected every later token of
//     cost >= c is rejected too.  Each precision class therefore keeps a PREFIX
//     of its tokens in walk order, i.e. class c is described by one threshold
//     key thr[c]: token kept  <=>  key(token) < thr[label(token)].
//   * key = (descending-order image of the score, index): ascending key order is
//     exactly the walk order with ties broken towards the lowest index.
//   * thresholds are found by a "bucketed walk": one histogram pass buckets the
//     keys (2048 range-compressed buckets, per class); buckets whose live tokens
//     all fit are taken whole by a block-wide prefix scan; only the <=3 buckets
//     where the budget runs out are materialised, sorted (tens of keys) and
//     walked literally.  Oversized boundary buckets are refined recursively.
// Costs are bits/8, so all budget arithmetic is exact in integer "bit" units:
//   cur + cost <= budget  <=>  8*cur + bits <= floor(8*budget).
#include "score_select.cuh"

namespace rtkv {

__global__ void __launch_bounds__(kT, 1) score_select_kernel(K2Params p)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    Smem &sm = *reinterpret_cast<Smem *>(dyn_smem);
    const int b = blockIdx.x, S = p.S;
    float *sc = p.scores + (int64_t)b * S;
    uint8_t *lb = p.labels + (int64_t)b * S;
    // staged copies of the row (scores + labels) in shared memory: every later pass reads these
    float *s_sc = reinterpret_cast<float *>(dyn_smem + ((sizeof(Smem) + 15) / 16) * 16);
    uint8_t *s_lb = reinterpret_cast<uint8_t *>(s_sc + ((S + 3) / 4) * 4);
    const float *rsc = p.staged ? s_sc : sc;
    const uint8_t *rlb = p.staged ? s_lb : lb;
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) { sm.dbg = p.dbg; sm.dbg_n = 0; tick(sm); }

    // ---- scores (token_importance.py:134-176) ----
    // Every loop below handles 4 independent tokens per trip (i, i+kT, i+2kT, i+3kT) with the
    // loads issued first: the per-token chain (L2 load -> 3 IEEE divisions -> stores) is long,
    // so instruction-level parallelism, not occupancy, is what keeps this single CTA moving.
    const float *__restrict__ ms = p.mass ? p.mass + (int64_t)b * S : nullptr;
    const float *__restrict__ tab = p.log_tab;
    float mn = 0.0f, den = 0.0f;
    if ((p.flags & RTKV_FROM_MASS) && p.mass_range) {
        const volatile unsigned *mr = p.mass_range + 4 * b;
        const unsigned kmin = mr[0], kmax = mr[1], nan = mr[2];
        __syncthreads();                                // everybody has read it ...
        if (threadIdx.x == 0) { unsigned *w = p.mass_range + 4 * b; w[0] = 0xffffffffu; w[1] = 0u; w[2] = 0u; }   // ... reset for the next layer
        mn = score_of_key(kmax);
        float mx = score_of_key(kmin);
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        den = rnd_dt(__fsub_rn(mx, mn), p.score_dt);
    } else if (p.flags & RTKV_FROM_MASS) {
        RowAgg a = agg_identity();                      // kmin/kmax hold the mass range as walk keys
        for (int i0 = threadIdx.x; i0 < S; i0 += 4 * kT) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * kT; v[u] = i < S ? ms[i] : ms[i0]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a.nan |= (v[u] != v[u]);
                const unsigned k = score_key_desc(v[u]);
                a.kmin = k < a.kmin ? k : a.kmin; a.kmax = k > a.kmax ? k : a.kmax;
            }
        }
        tick(sm);   // min/max loop done
        a = block_reduce_agg(a, sm, false);
        mn = score_of_key(a.kmax);                      // largest key = smallest value
        float mx = score_of_key(a.kmin);
        if (a.nan) mn = mx = __int_as_float(0x7fc00000);
        den = rnd_dt(__fsub_rn(mx, mn), p.score_dt);
    }
    tick(sm);   // 1: mass min/max done

    // ---- one pass: score, label, row statistics ----
    RowAgg st = agg_identity();
    {
        const int dt = p.score_dt;
        const bool ok = den > 1e-8f;
        const float t3 = __fmul_rn(p.gamma, p.rel);
        const bool from_mass = (p.flags & RTKV_FROM_MASS) != 0, make_labels = (p.flags & RTKV_MAKE_LABELS) != 0;
        for (int i0 = threadIdx.x; i0 < S; i0 += 4 * kT) {
            float sv[4], lg[4];
            int lv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {               // loads first
                const int i = i0 + u * kT;
                const bool in = i < S;
                if (from_mass) {
                    sv[u] = in ? ms[i] : 0.0f;
                    lg[u] = (in && i < p.log_tab_len) ? tab[i] : 0.0f;
                } else {
                    sv[u] = in ? sc[i] : 0.0f;
                }
                lv[u] = (!make_labels && in) ? lab_at(lb, i) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kT;
                if (i < S) {
                    float s = sv[u];
                    if (from_mass) {
                        float n = 0.0f, t1;
                        if (dt == RTKV_F32) {            // float32 attention: no per-op rounding to a half dtype
                            if (ok) n = __fdiv_rn(__fsub_rn(s, mn), den);
                            t1 = __fmul_rn(__fmul_rn(p.alpha, n), p.w_l);
                        } else {
                            if (ok) n = rnd_dt(__fdiv_rn(rnd_dt(__fsub_rn(s, mn), dt), den), dt);
                            t1 = rnd_dt(__fmul_rn(rnd_dt(__fmul_rn(p.alpha, n), dt), p.w_l), dt);
                        }
                        float pb = 0.0f;
                        if (S > 1) {
                            const float l = (i < p.log_tab_len) ? lg[u] : (float)log((double)(float)(i + 1));
                            pb = __fdiv_rn(l, p.log_S);
                        }
                        s = __fadd_rn(__fadd_rn(t1, __fmul_rn(p.beta, pb)), t3);
                        sc[i] = s;
                    }
                    int l = lv[u];
                    if (make_labels) { l = label_of(s, p.theta_h, p.theta_m); lb[i] = (uint8_t)l; }
                    if (p.staged) { s_sc[i] = s; s_lb[i] = (uint8_t)l; }
                    st.n0 += (l == 0); st.n1 += (l == 1); st.n2 += (l == 2);
                    st.sum += (double)s; st.sumsq += (double)s * (double)s;
                    st.nan |= (s != s);
                    const unsigned k = score_key_desc(s);
                    st.kmin = k < st.kmin ? k : st.kmin; st.kmax = k > st.kmax ? k : st.kmax;
                }
            }
        }
    }
    tick(sm);   // score loop done (before the statistics reduction)
    st = block_reduce_agg(st, sm);          // (also orders the staged writes before the reads below)
    float smin = score_of_key(st.kmax), smax = score_of_key(st.kmin);
    if (st.nan) smin = smax = __int_as_float(0x7fc00000);
    tick(sm);   // 2: scores, labels, statistics done

    // ---- budgeted selection ----
    int count = 0, sel_class[3] = {0, 0, 0};
    double sel_sum = 0.0;
    long long cost_bits = 0;
    if (p.flags & RTKV_SELECT) {
        SelCfg cfg;
        cfg.bits[0] = p.bits[0]; cfg.bits[1] = p.bits[1]; cfg.bits[2] = p.bits[2];
        cfg.budget_bits = p.budget_bits;
        const long long all_bits = (long long)cfg.bits[0] * st.n0 + (long long)cfg.bits[1] * st.n1 +
                                   (long long)cfg.bits[2] * st.n2;
        if (all_bits <= cfg.budget_bits) {            // the whole row fits: every token is kept, in order
            uint8_t *mask = p.mask + (int64_t)b * S;
            int32_t *sel_idx = p.sel_idx + (int64_t)b * S;
            for (int i = threadIdx.x; i < S; i += kT) { mask[i] = 1; sel_idx[i] = i; }
            cost_bits = all_bits; count = S; sel_sum = st.sum;
            sel_class[0] = st.n0; sel_class[1] = st.n1; sel_class[2] = st.n2;
            tick(sm);
        } else {
            select_thresholds(rsc, rlb, S, st.kmin, st.kmax, cfg, sm);
            cost_bits = sm.cur;
            tick(sm);   // thresholds done
            emit_selection(rsc, rlb, S, p.mask + (int64_t)b * S, p.sel_idx + (int64_t)b * S, sm, count,
                           sel_class, sel_sum);
        }
        tick(sm);   // emit done
    }
    if (threadIdx.x == 0) {
        rtkv_row_stats &r = p.row_stats[b];
        r.n_selected = count; r.n_selected_greedy = count;
        r.sel_class[0] = sel_class[0]; r.sel_class[1] = sel_class[1]; r.sel_class[2] = sel_class[2];
        r.label_count[0] = st.n0; r.label_count[1] = st.n1; r.label_count[2] = st.n2;
        r.cost_bits = cost_bits; r.sel_score_sum = sel_sum;
        r.score_sum = st.sum; r.score_sumsq = st.sumsq; r.score_min = smin; r.score_max = smax;
    }

    // ---- last CTA: layer statistics and the emergency fallback (selective_propagation.py:203-211)
    // The ticket lives in layer_stats.reserved[1]; it must be 0 on entry and is left 0.
    int mx = count;
    if (p.B > 1) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0)
            sm.bc_i = (atomicAdd((unsigned int *)&p.layer_stats->reserved[1], 1u) == (unsigned)(p.B - 1));
        __syncthreads();
        if (!sm.bc_i) return;
        __threadfence();
        mx = 0;
        for (int r = 0; r < p.B; ++r) {
            const int c = ((volatile rtkv_row_stats *)p.row_stats)[r].n_selected;
            mx = c > mx ? c : mx;
        }
    }
    int used_fb = 0;
    if ((p.flags & RTKV_SELECT) && (p.flags & RTKV_FALLBACK) && mx == 0 && p.fallback_topk > 0) {
        used_fb = 1;
        SelCfg cfg;
        cfg.bits[0] = cfg.bits[1] = cfg.bits[2] = 1;
        cfg.budget_bits = p.fallback_topk;
        for (int r = 0; r < p.B; ++r) {      // rare path: rows read from global memory
            const float *rs = p.scores + (int64_t)r * S;
            const uint8_t *rl = p.labels + (int64_t)r * S;
            RowAgg a = agg_identity();
            for (int i = threadIdx.x; i < S; i += kT) {
                const unsigned k = score_key_desc(rs[i]);
                a.kmin = k < a.kmin ? k : a.kmin; a.kmax = k > a.kmax ? k : a.kmax;
            }
            a = block_reduce_agg(a, sm, false);
            select_thresholds(rs, rl, S, a.kmin, a.kmax, cfg, sm);
            int cnt, cls[3]; double ss;
            emit_selection(rs, rl, S, p.mask + (int64_t)r * S, p.sel_idx + (int64_t)r * S, sm, cnt, cls, ss);
            if (threadIdx.x == 0) p.row_stats[r].n_selected = cnt;
            mx = cnt > mx ? cnt : mx;
        }
    }
    if (threadIdx.x == 0) {
        tick(sm);
        if (sm.dbg && blockIdx.x == 0) sm.dbg[0] = sm.dbg_n;
        p.layer_stats->max_selected = mx;
        p.layer_stats->used_fallback = used_fb;
        p.layer_stats->reserved[0] = 0;
        p.layer_stats->reserved[1] = 0;   // ticket back to zero for the next call
    }
}

// (float)log((double)(i+1)) table shared by every layer and every sequence length
__global__ void log_table_kernel(float *tab, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tab[i] = (float)log((double)(float)(i + 1));
}

// ---- small stand-alone kernels (component API) -----------------------------------
__global__ void position_bias_kernel(int S, float log_S, float *pb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) pb[i] = S > 1 ? __fdiv_rn((float)log((double)(float)(i + 1)), log_S) : 0.0f;
}

__global__ void __launch_bounds__(kT, 1) normalize_kernel(const float *mass, int dt, int S, float *out)
{
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    const float *ms = mass + (int64_t)blockIdx.x * S;
    float *o = out + (int64_t)blockIdx.x * S;
    float mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int i = threadIdx.x; i < S; i += kT) { float v = ms[i]; nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v); }
    mn = block_reduce(mn, OpMinF(), INFINITY, red_f);
    mx = block_reduce(mx, OpMaxF(), -INFINITY, red_f);
    nan = block_reduce(nan, OpAddI(), 0, red_i);
    if (nan) mn = mx = __int_as_float(0x7fc00000);
    const float den = rnd_dt(__fsub_rn(mx, mn), dt);
    for (int i = threadIdx.x; i < S; i += kT)
        o[i] = den > 1e-8f ? rnd_dt(__fdiv_rn(rnd_dt(__fsub_rn(ms[i], mn), dt), den), dt) : 0.0f;
}

__global__ void labels_to_i64_kernel(const uint8_t *l, int64_t n, int64_t *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = l[i];
}

__global__ void gather_scalars_kernel(const float *scores, const uint8_t *labels, const int32_t *sel_idx,
                                      const rtkv_row_stats *rs, int S, int out_rows,
                                      float *sel_scores, int64_t *sel_labels)
{
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= out_rows) return;
    float s = 0.0f; int64_t l = 0;
    if (r < rs[b].n_selected) {
        int i = sel_idx[(int64_t)b * S + r];
        s = scores[(int64_t)b * S + i];
        l = labels[(int64_t)b * S + i];
    }
    if (sel_scores) sel_scores[(int64_t)b * out_rows + r] = s;
    if (sel_labels) sel_labels[(int64_t)b * out_rows + r] = l;
}

} // namespace rtkv

using namespace rtkv;

static int g_k2_mode = 0;   // 0 auto (cluster from 8192 tokens), 1 single CTA, 2 cluster whenever S >= 1024
extern "C" int rtkv_debug_k2_mode(int mode)
{
    RTKV_CHECK_ARG(mode >= 0 && mode <= 2, "rtkv_debug_k2_mode: mode must be 0, 1 or 2");
    g_k2_mode = mode;
    return RTKV_OK;
}

static long long *g_k2_dbg = nullptr;
extern "C" int rtkv_debug_k2_timing(void *device_buffer_64_i64)
{
    g_k2_dbg = (long long *)device_buffer_64_i64;
    return RTKV_OK;
}

// scratch layout: [0,256) reserved words (whole-tensor min/max of rtkv_quant_params),
//                 [256, 256 + 4*n) log table of n = (bytes - 256) / 4 entries.
constexpr size_t kScratchHeader = 256;

extern "C" int rtkv_scratch_bytes(int B, int S, size_t *bytes)
{
    RTKV_CHECK_ARG(bytes && B > 0 && S > 0, "rtkv_scratch_bytes: bad argument");
    *bytes = kScratchHeader + 4 * (size_t)S;
    return RTKV_OK;
}

extern "C" int rtkv_scratch_init(void *scratch, size_t bytes, void *stream)
{
    RTKV_CHECK_ARG(scratch && bytes >= kScratchHeader, "rtkv_scratch_init: scratch too small");
    RTKV_CHECK_ARG((uintptr_t)scratch % 16 == 0, "rtkv_scratch_init: scratch must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = (bytes - kScratchHeader) / 4;
    const int n = (int)(cap < (size_t)1 << 30 ? cap : (size_t)1 << 30);
    RTKV_CUDA(cudaMemsetAsync(scratch, 0, kScratchHeader, st));
    if (n > 0) {
        log_table_kernel<<<(n + 255) / 256, 256, 0, st>>>((float *)((char *)scratch + kScratchHeader), n);
        RTKV_LAUNCH_CHECK("log_table_kernel");
    }
    return RTKV_OK;
}

static int score_select_impl(const float *mass, uint32_t *mass_range, float *scores, uint8_t *labels, uint8_t *mask,
                             int32_t *sel_idx, rtkv_row_stats *row_stats,
                             rtkv_layer_stats *layer_stats, const void *scratch, size_t scratch_bytes,
                             int B, int S, const rtkv_layer_params *prm, int flags, void *stream);

extern "C" int rtkv_score_select(const float *mass, float *scores, uint8_t *labels, uint8_t *mask,
                                 int32_t *sel_idx, rtkv_row_stats *row_stats,
                                 rtkv_layer_stats *layer_stats, const void *scratch, size_t scratch_bytes,
                                 int B, int S, const rtkv_layer_params *prm, int flags, void *stream)
{
    return score_select_impl(mass, nullptr, scores, labels, mask, sel_idx, row_stats, layer_stats, scratch, scratch_bytes, B, S, prm,
                             flags, stream);
}

extern "C" int rtkv_score_select_ranged(const float *mass, uint32_t *mass_range, float *scores, uint8_t *labels, uint8_t *mask,
                                        int32_t *sel_idx, rtkv_row_stats *row_stats,
                                        rtkv_layer_stats *layer_stats, const void *scratch, size_t scratch_bytes,
                                        int B, int S, const rtkv_layer_params *prm, int flags, void *stream)
{
    RTKV_CHECK_ARG(mass && mass_range && (flags & RTKV_FROM_MASS), "rtkv_score_select_ranged: needs mass, mass_range and RTKV_FROM_MASS");
    return score_select_impl(mass, mass_range, scores, labels, mask, sel_idx, row_stats, layer_stats, scratch, scratch_bytes, B, S,
                             prm, flags, stream);
}

static int score_select_impl(const float *mass, uint32_t *mass_range, float *scores, uint8_t *labels, uint8_t *mask,
                             int32_t *sel_idx, rtkv_row_stats *row_stats,
                             rtkv_layer_stats *layer_stats, const void *scratch, size_t scratch_bytes,
                             int B, int S, const rtkv_layer_params *prm, int flags, void *stream)
{
    RTKV_CHECK_ARG(prm, "rtkv_score_select: params is null");
    RTKV_CHECK_ARG(B > 0 && S > 0, "rtkv_score_select: bad shape B=%d S=%d", B, S);
    RTKV_CHECK_ARG(S <= (1 << 24), "rtkv_score_select: S > 2^24 is not supported");
    RTKV_CHECK_ARG(scores && labels && row_stats && layer_stats, "rtkv_score_select: null pointer");
    RTKV_CHECK_ARG(!(flags & RTKV_FROM_MASS) || mass, "rtkv_score_select: mass is null");
    RTKV_CHECK_ARG(!(flags & RTKV_SELECT) || (mask && sel_idx), "rtkv_score_select: mask/sel_idx is null");
    RTKV_CHECK_ARG(dtype_ok(prm->score_dtype), "rtkv_score_select: bad score_dtype");
    if (flags & RTKV_SELECT)
        for (int c = 0; c < 3; ++c)
            RTKV_CHECK_ARG(prm->bits[c] >= 1 && prm->bits[c] <= 24, "rtkv_score_select: bits[%d]=%d outside 1..24", c, prm->bits[c]);
    K2Params p;
    p.mass = mass; p.mass_range = mass_range; p.scores = scores; p.labels = labels; p.mask = mask; p.sel_idx = sel_idx;
    p.row_stats = row_stats; p.layer_stats = layer_stats;
    p.dbg = g_k2_dbg;
    p.log_tab = nullptr; p.log_tab_len = 0;
    if (scratch && scratch_bytes > kScratchHeader) {   // table length is implied by the size rtkv_scratch_init saw
        p.log_tab = (const float *)((const char *)scratch + kScratchHeader);
        p.log_tab_len = (int)((scratch_bytes - kScratchHeader) / 4);
    }
    p.B = B; p.S = S; p.flags = flags; p.score_dt = prm->score_dtype;
    p.alpha = (float)prm->alpha; p.beta = (float)prm->beta; p.gamma = (float)prm->gamma;
    p.w_l = (float)prm->layer_weight; p.theta_h = (float)prm->theta_h; p.theta_m = (float)prm->theta_m;
    double rel = (double)prm->prompt_len / (double)S;                 // token_importance.py:129
    p.rel = (float)(rel < 1.0 ? rel : 1.0);
    p.log_S = (float)log((double)S);                                  // token_importance.py:108
    for (int c = 0; c < 3; ++c) p.bits[c] = prm->bits[c];
    double b8 = prm->budget * 8.0;
    if (!(b8 >= 0.0)) p.budget_bits = -1;                              // negative or NaN: nothing fits
    else if (b8 >= 9.0e18) p.budget_bits = LLONG_MAX / 2;
    else p.budget_bits = (long long)floor(b8);
    p.fallback_topk = prm->fallback_topk < S ? prm->fallback_topk : S;
    if (g_k2_mode != 1) {   // one row across a cluster of 8 CTAs (8 SMs) when it applies
        const int rc = launch_score_select_cluster(p, (cudaStream_t)stream, g_k2_mode == 2);
        if (rc != RTKV_EUNSUPPORTED) return rc;
        // mode 2 with a row too short (or too long) for the cluster: fall through to the single-CTA kernel
    }
    // shared memory: control block + (when it fits) the row's scores and labels
    const size_t base = ((sizeof(Smem) + 15) / 16) * 16;
    const size_t staged = base + 4 * (size_t)((S + 3) / 4 * 4) + (size_t)S;
    p.staged = staged <= 200 * 1024;
    const size_t smem = p.staged ? staged : base;
    static size_t attr_bytes[64] = {0};
    int dev = 0;
    RTKV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_bytes[dev] < smem) {
        RTKV_CUDA(cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_bytes[dev] = 200 * 1024;
    }
    RTKV_CUDA(launch_kernel(score_select_kernel, dim3(B), dim3(kT), smem, (cudaStream_t)stream, 1, p));
    return RTKV_OK;
}

extern "C" int rtkv_position_bias(int S, float *pb, void *stream)
{
    RTKV_CHECK_ARG(S >= 0 && (pb || S == 0), "rtkv_position_bias: bad argument");
    if (S == 0) return RTKV_OK;
    position_bias_kernel<<<(S + 255) / 256, 256, 0, (cudaStream_t)stream>>>(S, (float)log((double)S), pb);
    RTKV_LAUNCH_CHECK("position_bias_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_normalize_mass(const float *mass, int dtype, int B, int S, float *out, void *stream)
{
    RTKV_CHECK_ARG(mass && out && B > 0 && S > 0 && dtype_ok(dtype), "rtkv_normalize_mass: bad argument");
    normalize_kernel<<<B, kT, 0, (cudaStream_t)stream>>>(mass, dtype, S, out);
    RTKV_LAUNCH_CHECK("normalize_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_labels_to_i64(const uint8_t *labels, int64_t n, int64_t *out, void *stream)
{
    RTKV_CHECK_ARG(labels && out && n >= 0, "rtkv_labels_to_i64: bad argument");
    if (n == 0) return RTKV_OK;
    labels_to_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, n, out);
    RTKV_LAUNCH_CHECK("labels_to_i64_kernel");
    return RTKV_OK;
}

extern "C" int rtkv_gather_scalars(const float *scores, const uint8_t *labels, const int32_t *sel_idx,
                                   const rtkv_row_stats *row_stats, int B, int S, int out_rows,
                                   float *sel_scores, int64_t *sel_labels, void *stream)
{
    RTKV_CHECK_ARG(scores && labels && sel_idx && row_stats && B > 0 && S > 0 && out_rows >= 0,
                   "rtkv_gather_scalars: bad argument");
    if (out_rows == 0) return RTKV_OK;
    dim3 grid((out_rows + 255) / 256, B);
    gather_scalars_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, labels, sel_idx, row_stats, S,
                                                                 out_rows, sel_scores, sel_labels);
    RTKV_LAUNCH_CHECK("gather_scalars_kernel");
    return RTKV_OK;
}
