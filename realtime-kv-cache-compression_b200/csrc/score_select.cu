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
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace rtkv {

constexpr int kT = 1024;        // threads per row CTA
constexpr int kNB = 2048;       // buckets per refinement level (2 per thread)
constexpr int kMaxCand = 1024;  // boundary-bucket tokens sorted in shared memory
constexpr int kMaxDepth = 8;

struct K2Params {
    const float *mass;
    float *scores;
    uint8_t *labels;
    uint8_t *mask;
    int32_t *sel_idx;
    rtkv_row_stats *row_stats;
    rtkv_layer_stats *layer_stats;
    unsigned int *ticket;
    int B, S, flags, score_dt;
    float alpha, beta, gamma, w_l, theta_h, theta_m, rel, log_S;
    int bits[3];
    long long budget_bits;
    int fallback_topk;
};

struct Smem {
    int hist[3][kNB];
    unsigned long long cand_key[kMaxCand];
    unsigned long long sort_key[kMaxCand];
    uint8_t cand_lab[kMaxCand];
    uint8_t sort_lab[kMaxCand];
    long long red_ll[32];
    double red_d[32];
    float red_f[32];
    int red_i[32];
    unsigned red_u[32];
    unsigned long long thr[3];
    unsigned long long lvl_lo[kMaxDepth];
    int lvl_shift[kMaxDepth], lvl_nb[kMaxDepth], lvl_resume[kMaxDepth];
    long long cur, sum_before, total;
    int alive[3];
    int depth, found, ncand, run_base, bc_i;
};

// ---- block reductions (all 1024 threads; result broadcast; fixed order) ----------
template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T ident, T *buf)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) buf[warp] = v;
    __syncthreads();
    T r = (lane < (int)(blockDim.x >> 5)) ? buf[lane] : ident;
#pragma unroll
    for (int o = 16; o; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    __syncthreads();
    return r;
}
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpAddI { __device__ int operator()(int a, int b) const { return a + b; } };
struct OpMinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinU { __device__ unsigned operator()(unsigned a, unsigned b) const { return a < b ? a : b; } };
struct OpMaxU { __device__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; } };

// ---- walk-order key ------------------------------------------------------------------
// torch.argsort(descending=True): NaN first, then +inf ... -inf, -0 == +0.
__device__ __forceinline__ unsigned score_key_desc(float s)
{
    if (s != s) return 0u;
    unsigned u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;
    unsigned asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}
__device__ __forceinline__ unsigned long long walk_key(float s, int i)
{
    return ((unsigned long long)score_key_desc(s) << 32) | (unsigned)i;
}

__device__ __forceinline__ uint8_t label_of(float s, float th, float tm)
{
    // high = s >= th ; medium = (s >= tm) & (s < th)   (dynamic_quantization.py:41-45)
    return (s >= th) ? 2 : (((s >= tm) && (s < th)) ? 1 : 0);
}

struct SelCfg { int bits[3]; long long budget_bits; };

// caller-supplied labels are clamped so that a bad value cannot index out of bounds
__device__ __forceinline__ int lab_at(const uint8_t *lb, int i) { int l = lb[i]; return l > 2 ? 2 : l; }

// Histogram of the current level: tokens whose key lies in the level's range, per class.
__device__ void build_hist(const float *sc, const uint8_t *lb, int S, Smem &sm)
{
    const int d = sm.depth;
    const unsigned long long lo = sm.lvl_lo[d];
    const int shift = sm.lvl_shift[d], nb = sm.lvl_nb[d];
    for (int x = threadIdx.x; x < 3 * kNB; x += kT) (&sm.hist[0][0])[x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += kT) {
        unsigned long long k = walk_key(sc[i], i);
        if (k >= lo) {
            unsigned long long bk = (k - lo) >> shift;
            if (bk < (unsigned long long)nb) atomicAdd(&sm.hist[lab_at(lb, i)][(int)bk], 1);
        }
    }
    __syncthreads();
}

// First bucket x >= b0 where the live tokens up to and including x no longer fit.
// Leaves sm.found (INT_MAX if everything fits), sm.sum_before, sm.total.
__device__ void scan_find(int b0, const SelCfg &cfg, Smem &sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = sm.lvl_nb[sm.depth];
    const long long w0c = sm.alive[0] ? cfg.bits[0] : 0, w1c = sm.alive[1] ? cfg.bits[1] : 0,
                    w2c = sm.alive[2] ? cfg.bits[2] : 0;
    const long long room = cfg.budget_bits - sm.cur;
    long long w[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int x = 2 * threadIdx.x + k;
        w[k] = (x >= b0 && x < nb)
                   ? w0c * sm.hist[0][x] + w1c * sm.hist[1][x] + w2c * sm.hist[2][x] : 0;
    }
    long long incl = w[0] + w[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.red_ll[warp] = incl;
    if (threadIdx.x == 0) sm.found = INT_MAX;
    __syncthreads();
    long long woff = 0, tot = 0;
    for (int k = 0; k < 32; ++k) {
        long long t = sm.red_ll[k];
        if (k < warp) woff += t;
        tot += t;
    }
    const long long excl = woff + incl - (w[0] + w[1]);
    const long long p0 = excl + w[0], p1 = p0 + w[1];
    int mine = INT_MAX;
    if (w[0] > 0 && p0 > room) mine = 2 * threadIdx.x;
    else if (w[1] > 0 && p1 > room) mine = 2 * threadIdx.x + 1;
    if (mine != INT_MAX) atomicMin(&sm.found, mine);
    __syncthreads();
    const int f = sm.found;
    if (f != INT_MAX && (f >> 1) == (int)threadIdx.x) sm.sum_before = (f & 1) ? p0 : excl;
    if (threadIdx.x == 0) sm.total = tot;
    __syncthreads();
}

// Materialise the live tokens of bucket bsel, sort them by key and walk them literally.
__device__ void walk_bucket(const float *sc, const uint8_t *lb, int S, int bsel,
                            const SelCfg &cfg, Smem &sm)
{
    const int d = sm.depth;
    const unsigned long long lo = sm.lvl_lo[d];
    const int shift = sm.lvl_shift[d];
    if (threadIdx.x == 0) sm.ncand = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += kT) {
        unsigned long long k = walk_key(sc[i], i);
        if (k >= lo && ((k - lo) >> shift) == (unsigned long long)bsel) {
            int l = lab_at(lb, i);
            if (sm.alive[l]) {
                int pos = atomicAdd(&sm.ncand, 1);
                if (pos < kMaxCand) { sm.cand_key[pos] = k; sm.cand_lab[pos] = (uint8_t)l; }
            }
        }
    }
    __syncthreads();
    const int m = sm.ncand < kMaxCand ? sm.ncand : kMaxCand;
    if ((int)threadIdx.x < m) { // rank by counting: keys are unique
        const unsigned long long k = sm.cand_key[threadIdx.x];
        int rank = 0;
        for (int u = 0; u < m; ++u) rank += sm.cand_key[u] < k;
        sm.sort_key[rank] = k;
        sm.sort_lab[rank] = sm.cand_lab[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long cur = sm.cur;
        for (int t = 0; t < m; ++t) {
            const int c = sm.sort_lab[t];
            if (!sm.alive[c]) continue;
            if (cur + cfg.bits[c] <= cfg.budget_bits) cur += cfg.bits[c];
            else {
                for (int c2 = 0; c2 < 3; ++c2)
                    if (sm.alive[c2] && cfg.bits[c2] >= cfg.bits[c]) {
                        sm.alive[c2] = 0;
                        sm.thr[c2] = sm.sort_key[t];
                    }
            }
        }
        sm.cur = cur;
    }
    __syncthreads();
}

// Thresholds thr[0..2] of one row.  kmin/kmax: range of score_key_desc over the row.
__device__ void select_thresholds(const float *sc, const uint8_t *lb, int S, unsigned kmin,
                                  unsigned kmax, const SelCfg &cfg, Smem &sm)
{
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) { sm.alive[c] = 1; sm.thr[c] = ~0ull; }
        sm.cur = 0;
        sm.depth = 0;
        const unsigned long long lo = (unsigned long long)kmin << 32;
        const unsigned long long span = (((unsigned long long)kmax << 32) | (unsigned)(S - 1)) - lo;
        int shift = span ? (64 - __clzll((long long)span)) - 11 : 0;
        if (shift < 0) shift = 0;
        sm.lvl_lo[0] = lo; sm.lvl_shift[0] = shift; sm.lvl_nb[0] = (int)(span >> shift) + 1;
        sm.lvl_resume[0] = 0;
    }
    __syncthreads();
    bool need_hist = true;
    for (;;) {
        if (need_hist) build_hist(sc, lb, S, sm);
        need_hist = false;
        if (!(sm.alive[0] | sm.alive[1] | sm.alive[2])) break;
        const int d = sm.depth;
        scan_find(sm.lvl_resume[d], cfg, sm);
        const int f = sm.found;
        if (f == INT_MAX) {                 // every live token of this level fits
            __syncthreads();
            if (threadIdx.x == 0) { sm.cur += sm.total; if (d > 0) sm.depth = d - 1; }
            __syncthreads();
            if (d == 0) break;
            need_hist = true;               // back in the parent level, after the refined bucket
            continue;
        }
        const int m_alive = (sm.alive[0] ? sm.hist[0][f] : 0) + (sm.alive[1] ? sm.hist[1][f] : 0) +
                            (sm.alive[2] ? sm.hist[2][f] : 0);
        const int shift = sm.lvl_shift[d];
        __syncthreads();
        if (threadIdx.x == 0) { sm.cur += sm.sum_before; sm.lvl_resume[d] = f + 1; }
        __syncthreads();
        if (m_alive <= kMaxCand || shift == 0 || d + 1 >= kMaxDepth) {
            walk_bucket(sc, lb, S, f, cfg, sm);
        } else {                            // refine the boundary bucket
            if (threadIdx.x == 0) {
                const int ns = shift > 11 ? shift - 11 : 0;
                sm.lvl_lo[d + 1] = sm.lvl_lo[d] + ((unsigned long long)f << shift);
                sm.lvl_shift[d + 1] = ns;
                sm.lvl_nb[d + 1] = 1 << (shift - ns);
                sm.lvl_resume[d + 1] = 0;
                sm.depth = d + 1;
            }
            __syncthreads();
            need_hist = true;
        }
    }
    __syncthreads();
}

// mask, compacted ascending index list and the per-row selection statistics
__device__ void emit_selection(const float *sc, const uint8_t *lb, int S, uint8_t *mask,
                               int32_t *sel_idx, Smem &sm, int &count, int sel_class[3],
                               double &sel_sum)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long t0 = sm.thr[0], t1 = sm.thr[1], t2 = sm.thr[2];
    int base = 0, c0 = 0, c1 = 0, c2 = 0;
    double ssum = 0.0;
    for (int tile = 0; tile < S; tile += kT) {
        const int i = tile + threadIdx.x;
        bool keep = false;
        float s = 0.0f;
        int l = 0;
        if (i < S) {
            s = sc[i]; l = lab_at(lb, i);
            const unsigned long long k = walk_key(s, i);
            keep = k < (l == 0 ? t0 : (l == 1 ? t1 : t2));
            mask[i] = keep ? 1 : 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sm.red_i[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int k = 0; k < 32; ++k) { int t = sm.red_i[k]; if (k < warp) woff += t; tot += t; }
        if (keep) {
            sel_idx[base + woff + __popc(bal & ((1u << lane) - 1u))] = i;
            ssum += (double)s;
            c0 += (l == 0); c1 += (l == 1); c2 += (l == 2);
        }
        base += tot;
        __syncthreads();
    }
    count = base;
    sel_sum = block_reduce(ssum, OpAddD(), 0.0, sm.red_d);
    sel_class[0] = block_reduce(c0, OpAddI(), 0, sm.red_i);
    sel_class[1] = block_reduce(c1, OpAddI(), 0, sm.red_i);
    sel_class[2] = block_reduce(c2, OpAddI(), 0, sm.red_i);
}

__global__ void __launch_bounds__(kT, 1) score_select_kernel(K2Params p)
{
    __shared__ Smem sm;
    const int b = blockIdx.x, S = p.S;
    float *sc = p.scores + (int64_t)b * S;
    uint8_t *lb = p.labels + (int64_t)b * S;

    // ---- scores (token_importance.py:134-176) ----
    if (p.flags & RTKV_FROM_MASS) {
        const float *ms = p.mass + (int64_t)b * S;
        float mn = INFINITY, mx = -INFINITY;
        int nan = 0;
        for (int i = threadIdx.x; i < S; i += kT) {
            float v = ms[i];
            nan |= (v != v);
            mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
        mn = block_reduce(mn, OpMinF(), INFINITY, sm.red_f);
        mx = block_reduce(mx, OpMaxF(), -INFINITY, sm.red_f);
        nan = block_reduce(nan, OpAddI(), 0, sm.red_i);
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        const int dt = p.score_dt;
        const float den = rnd_dt(__fsub_rn(mx, mn), dt);
        const bool ok = den > 1e-8f;
        const float t3 = __fmul_rn(p.gamma, p.rel);
        for (int i = threadIdx.x; i < S; i += kT) {
            float n = 0.0f;
            if (ok) n = rnd_dt(__fdiv_rn(rnd_dt(__fsub_rn(ms[i], mn), dt), den), dt);
            const float t1 = rnd_dt(__fmul_rn(rnd_dt(__fmul_rn(p.alpha, n), dt), p.w_l), dt);
            float pb = 0.0f;
            if (S > 1) pb = __fdiv_rn((float)log((double)(float)(i + 1)), p.log_S);
            const float t2 = __fmul_rn(p.beta, pb);
            sc[i] = __fadd_rn(__fadd_rn(t1, t2), t3);
        }
        __syncthreads();
    }

    // ---- labels + row statistics ----
    double sum = 0.0, sumsq = 0.0;
    float smin = INFINITY, smax = -INFINITY;
    int n0 = 0, n1 = 0, n2 = 0, snan = 0;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = threadIdx.x; i < S; i += kT) {
        const float s = sc[i];
        int l;
        if (p.flags & RTKV_MAKE_LABELS) { l = label_of(s, p.theta_h, p.theta_m); lb[i] = (uint8_t)l; }
        else l = lab_at(lb, i);
        n0 += (l == 0); n1 += (l == 1); n2 += (l == 2);
        sum += (double)s; sumsq += (double)s * (double)s;
        snan |= (s != s);
        smin = fminf(smin, s); smax = fmaxf(smax, s);
        const unsigned k = score_key_desc(s);
        kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax;
    }
    sum = block_reduce(sum, OpAddD(), 0.0, sm.red_d);
    sumsq = block_reduce(sumsq, OpAddD(), 0.0, sm.red_d);
    smin = block_reduce(smin, OpMinF(), INFINITY, sm.red_f);
    smax = block_reduce(smax, OpMaxF(), -INFINITY, sm.red_f);
    snan = block_reduce(snan, OpAddI(), 0, sm.red_i);
    n0 = block_reduce(n0, OpAddI(), 0, sm.red_i);
    n1 = block_reduce(n1, OpAddI(), 0, sm.red_i);
    n2 = block_reduce(n2, OpAddI(), 0, sm.red_i);
    kmin = block_reduce(kmin, OpMinU(), 0xffffffffu, sm.red_u);
    kmax = block_reduce(kmax, OpMaxU(), 0u, sm.red_u);
    if (snan) smin = smax = __int_as_float(0x7fc00000);

    // ---- budgeted selection ----
    int count = 0, sel_class[3] = {0, 0, 0};
    double sel_sum = 0.0;
    long long cost_bits = 0;
    if (p.flags & RTKV_SELECT) {
        SelCfg cfg;
        cfg.bits[0] = p.bits[0]; cfg.bits[1] = p.bits[1]; cfg.bits[2] = p.bits[2];
        cfg.budget_bits = p.budget_bits;
        select_thresholds(sc, lb, S, kmin, kmax, cfg, sm);
        cost_bits = sm.cur;
        emit_selection(sc, lb, S, p.mask + (int64_t)b * S, p.sel_idx + (int64_t)b * S, sm, count,
                       sel_class, sel_sum);
    }
    if (threadIdx.x == 0) {
        rtkv_row_stats &r = p.row_stats[b];
        r.n_selected = count; r.n_selected_greedy = count;
        r.sel_class[0] = sel_class[0]; r.sel_class[1] = sel_class[1]; r.sel_class[2] = sel_class[2];
        r.label_count[0] = n0; r.label_count[1] = n1; r.label_count[2] = n2;
        r.cost_bits = cost_bits; r.sel_score_sum = sel_sum;
        r.score_sum = sum; r.score_sumsq = sumsq; r.score_min = smin; r.score_max = smax;
    }

    // ---- last CTA: layer statistics and the emergency fallback (selective_propagation.py:203-211)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sm.bc_i = (atomicAdd(p.ticket, 1u) == (unsigned)(p.B - 1));
    __syncthreads();
    if (!sm.bc_i) return;
    __threadfence();
    int mx = 0;
    for (int r = 0; r < p.B; ++r) { int c = ((volatile rtkv_row_stats *)p.row_stats)[r].n_selected; mx = c > mx ? c : mx; }
    int used_fb = 0;
    if ((p.flags & RTKV_SELECT) && (p.flags & RTKV_FALLBACK) && mx == 0 && p.fallback_topk > 0) {
        used_fb = 1;
        SelCfg cfg;
        cfg.bits[0] = cfg.bits[1] = cfg.bits[2] = 1;
        cfg.budget_bits = p.fallback_topk;
        for (int r = 0; r < p.B; ++r) {
            const float *rs = p.scores + (int64_t)r * S;
            const uint8_t *rl = p.labels + (int64_t)r * S;
            unsigned a = 0xffffffffu, z = 0u;
            for (int i = threadIdx.x; i < S; i += kT) {
                const unsigned k = score_key_desc(rs[i]);
                a = k < a ? k : a; z = k > z ? k : z;
            }
            a = block_reduce(a, OpMinU(), 0xffffffffu, sm.red_u);
            z = block_reduce(z, OpMaxU(), 0u, sm.red_u);
            select_thresholds(rs, rl, S, a, z, cfg, sm);
            int cnt, cls[3]; double ss;
            emit_selection(rs, rl, S, p.mask + (int64_t)r * S, p.sel_idx + (int64_t)r * S, sm, cnt, cls, ss);
            if (threadIdx.x == 0) p.row_stats[r].n_selected = cnt;
            mx = cnt > mx ? cnt : mx;
        }
    }
    if (threadIdx.x == 0) {
        p.layer_stats->max_selected = mx;
        p.layer_stats->used_fallback = used_fb;
        p.layer_stats->reserved[0] = p.layer_stats->reserved[1] = 0;
        *p.ticket = 0u; // leave the scratch zeroed for the next call
    }
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

extern "C" int rtkv_scratch_bytes(int B, int S, size_t *bytes)
{
    RTKV_CHECK_ARG(bytes && B > 0 && S > 0, "rtkv_scratch_bytes: bad argument");
    *bytes = 256;
    return RTKV_OK;
}

extern "C" int rtkv_score_select(const float *mass, float *scores, uint8_t *labels, uint8_t *mask,
                                 int32_t *sel_idx, rtkv_row_stats *row_stats,
                                 rtkv_layer_stats *layer_stats, void *scratch, int B, int S,
                                 const rtkv_layer_params *prm, int flags, void *stream)
{
    RTKV_CHECK_ARG(prm, "rtkv_score_select: params is null");
    RTKV_CHECK_ARG(B > 0 && S > 0, "rtkv_score_select: bad shape B=%d S=%d", B, S);
    RTKV_CHECK_ARG(S < (1 << 30), "rtkv_score_select: S too large");
    RTKV_CHECK_ARG(scores && labels && row_stats && layer_stats && scratch, "rtkv_score_select: null pointer");
    RTKV_CHECK_ARG(!(flags & RTKV_FROM_MASS) || mass, "rtkv_score_select: mass is null");
    RTKV_CHECK_ARG(!(flags & RTKV_SELECT) || (mask && sel_idx), "rtkv_score_select: mask/sel_idx is null");
    RTKV_CHECK_ARG(dtype_ok(prm->score_dtype), "rtkv_score_select: bad score_dtype");
    if (flags & RTKV_SELECT)
        for (int c = 0; c < 3; ++c)
            RTKV_CHECK_ARG(prm->bits[c] >= 1 && prm->bits[c] <= 24, "rtkv_score_select: bits[%d]=%d outside 1..24", c, prm->bits[c]);
    K2Params p;
    p.mass = mass; p.scores = scores; p.labels = labels; p.mask = mask; p.sel_idx = sel_idx;
    p.row_stats = row_stats; p.layer_stats = layer_stats; p.ticket = (unsigned int *)scratch;
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
    score_select_kernel<<<B, kT, 0, (cudaStream_t)stream>>>(p);
    RTKV_LAUNCH_CHECK("score_select_kernel");
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
