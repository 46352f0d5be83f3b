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
    long long *dbg;
    const float *log_tab;   // (float)log((double)(i+1)), i < log_tab_len; NULL: compute in place
    int log_tab_len;
    int staged;             // scores/labels of the row are staged in shared memory
    float *scores;
    uint8_t *labels;
    uint8_t *mask;
    int32_t *sel_idx;
    rtkv_row_stats *row_stats;
    rtkv_layer_stats *layer_stats;
    int B, S, flags, score_dt;
    float alpha, beta, gamma, w_l, theta_h, theta_m, rel, log_S;
    int bits[3];
    long long budget_bits;
    int fallback_topk;
};

struct Smem {
    long long *dbg;          // optional phase timestamps (rtkv_debug_k2_timing)
    int dbg_n;
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
    double agg_d[33][2];
    unsigned agg_u[33][2];
    int agg_i[33][4];
    int scan_pre[32];
    int scan_total;
    unsigned long long thr[3];
    unsigned long long lvl_lo[kMaxDepth];
    int lvl_shift[kMaxDepth], lvl_nb[kMaxDepth], lvl_resume[kMaxDepth];
    long long cur, sum_before, total;
    int alive[3];
    int depth, found, ncand, run_base, bc_i;
};

// phase timestamps for profiling: dbg[0] = count, dbg[1..] = clock64() at each tick
__device__ __forceinline__ void tick(Smem &sm)
{
    if (sm.dbg && threadIdx.x == 0 && blockIdx.x == 0 && sm.dbg_n < 62) sm.dbg[1 + sm.dbg_n++] = clock64();
}

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

// One fused block reduction for a whole record of row statistics: integer fields go
// through the single-instruction warp reductions (REDUX), the two double sums through
// shuffles; only warp 0 combines the 32 per-warp records (2 barriers in total).
// Score min/max are carried as walk keys: kmin <-> largest score, kmax <-> smallest.
struct RowAgg {
    double sum, sumsq;
    unsigned kmin, kmax;
    int n0, n1, n2, nan;
};
__device__ __forceinline__ RowAgg agg_identity()
{
    RowAgg a;
    a.sum = a.sumsq = 0.0; a.kmin = 0xffffffffu; a.kmax = 0u;
    a.n0 = a.n1 = a.n2 = a.nan = 0;
    return a;
}
__device__ __forceinline__ RowAgg agg_warp(RowAgg a, bool with_sums)
{
    a.kmin = __reduce_min_sync(0xffffffffu, a.kmin);
    a.kmax = __reduce_max_sync(0xffffffffu, a.kmax);
    a.n0 = __reduce_add_sync(0xffffffffu, a.n0);
    a.n1 = __reduce_add_sync(0xffffffffu, a.n1);
    a.n2 = __reduce_add_sync(0xffffffffu, a.n2);
    a.nan = __reduce_or_sync(0xffffffffu, (unsigned)a.nan);
    if (with_sums) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a.sum += __shfl_xor_sync(0xffffffffu, a.sum, o);
            a.sumsq += __shfl_xor_sync(0xffffffffu, a.sumsq, o);
        }
    }
    return a;
}
__device__ RowAgg block_reduce_agg(RowAgg a, Smem &sm, bool with_sums = true)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = agg_warp(a, with_sums);
    if (lane == 0) {
        sm.agg_d[warp][0] = a.sum; sm.agg_d[warp][1] = a.sumsq;
        sm.agg_u[warp][0] = a.kmin; sm.agg_u[warp][1] = a.kmax;
        sm.agg_i[warp][0] = a.n0; sm.agg_i[warp][1] = a.n1; sm.agg_i[warp][2] = a.n2; sm.agg_i[warp][3] = a.nan;
    }
    __syncthreads();
    if (warp == 0) {
        RowAgg r;
        r.sum = sm.agg_d[lane][0]; r.sumsq = sm.agg_d[lane][1];
        r.kmin = sm.agg_u[lane][0]; r.kmax = sm.agg_u[lane][1];
        r.n0 = sm.agg_i[lane][0]; r.n1 = sm.agg_i[lane][1]; r.n2 = sm.agg_i[lane][2]; r.nan = sm.agg_i[lane][3];
        r = agg_warp(r, with_sums);
        if (lane == 0) {
            sm.agg_d[32][0] = r.sum; sm.agg_d[32][1] = r.sumsq;
            sm.agg_u[32][0] = r.kmin; sm.agg_u[32][1] = r.kmax;
            sm.agg_i[32][0] = r.n0; sm.agg_i[32][1] = r.n1; sm.agg_i[32][2] = r.n2; sm.agg_i[32][3] = r.nan;
        }
    }
    __syncthreads();
    RowAgg r;
    r.sum = sm.agg_d[32][0]; r.sumsq = sm.agg_d[32][1];
    r.kmin = sm.agg_u[32][0]; r.kmax = sm.agg_u[32][1];
    r.n0 = sm.agg_i[32][0]; r.n1 = sm.agg_i[32][1]; r.n2 = sm.agg_i[32][2]; r.nan = sm.agg_i[32][3];
    return r;
}

// block-wide exclusive scan of one int per thread (+ total); only warp 0 scans the warp totals
__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ int block_excl_scan(int v, int &total, Smem &sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v, lane);
    if (lane == 31) sm.red_i[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int t = sm.red_i[lane];
        const int ti = warp_incl_scan(t, lane);
        sm.scan_pre[lane] = ti - t;
        if (lane == 31) sm.scan_total = ti;
    }
    __syncthreads();
    total = sm.scan_total;
    return sm.scan_pre[warp] + incl - v;
}

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
// inverse of score_key_desc (NaN is reported through the nan flag, -0 comes back as +0)
__device__ __forceinline__ float score_of_key(unsigned k)
{
    const unsigned asc = ~k;
    return __uint_as_float((asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc);
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
#pragma unroll 4
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
// (weights fit in 32 bits: S <= 2^24 tokens x <= 24 bits)
__device__ void scan_find(int b0, const SelCfg &cfg, Smem &sm)
{
    const int nb = sm.lvl_nb[sm.depth];
    const int w0c = sm.alive[0] ? cfg.bits[0] : 0, w1c = sm.alive[1] ? cfg.bits[1] : 0,
              w2c = sm.alive[2] ? cfg.bits[2] : 0;
    const long long room_ll = cfg.budget_bits - sm.cur;
    const int room = room_ll > 0x7fffffffLL ? 0x7fffffff : (room_ll < -1 ? -1 : (int)room_ll);
    int w[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = 2 * threadIdx.x + k;
        w[k] = (x >= b0 && x < nb) ? w0c * sm.hist[0][x] + w1c * sm.hist[1][x] + w2c * sm.hist[2][x] : 0;
    }
    if (threadIdx.x == 0) sm.found = INT_MAX;
    int tot;
    const int excl = block_excl_scan(w[0] + w[1], tot, sm);   // (barriers also order the sm.found reset)
    const int p0 = excl + w[0], p1 = p0 + w[1];
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
#pragma unroll 4
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
        if (need_hist) { build_hist(sc, lb, S, sm); tick(sm); }
        need_hist = false;
        if (!(sm.alive[0] | sm.alive[1] | sm.alive[2])) break;
        const int d = sm.depth;
        scan_find(sm.lvl_resume[d], cfg, sm);
        tick(sm);
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
            tick(sm);
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

// mask, compacted ascending index list and the per-row selection statistics.
// Warp w owns the contiguous token range [w*chunk, (w+1)*chunk): it counts its kept tokens,
// one block-wide scan turns the 32 counts into output offsets, then it writes its part.
__device__ void emit_selection(const float *sc, const uint8_t *lb, int S, uint8_t *mask,
                               int32_t *sel_idx, Smem &sm, int &count, int sel_class[3],
                               double &sel_sum)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long t0 = sm.thr[0], t1 = sm.thr[1], t2 = sm.thr[2];
    const int chunk = ((S + 31) / 32 + 127) / 128 * 128;        // multiple of 128 tokens per warp
    const int beg = warp * chunk, end = min(S, beg + chunk);
    int mine = 0;
    for (int base = beg; base < end; base += 128) {
        float sv[4];
        int lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            sv[u] = i < end ? sc[i] : 0.0f;
            lv[u] = i < end ? lab_at(lb, i) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            const bool keep = i < end && walk_key(sv[u], i) < (lv[u] == 0 ? t0 : (lv[u] == 1 ? t1 : t2));
            mine += __popc(__ballot_sync(0xffffffffu, keep));
        }
    }
    tick(sm);   // emit: counting pass done
    int total;
    int off = block_excl_scan(lane == 0 ? mine : 0, total, sm);
    off = __shfl_sync(0xffffffffu, off, 0);
    tick(sm);   // emit: scan done
    RowAgg a = agg_identity();
    for (int base = beg; base < end; base += 128) {
        float sv[4];
        int lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            sv[u] = i < end ? sc[i] : 0.0f;
            lv[u] = i < end ? lab_at(lb, i) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            const int l = lv[u];
            const bool keep = i < end && walk_key(sv[u], i) < (l == 0 ? t0 : (l == 1 ? t1 : t2));
            if (i < end) mask[i] = keep ? 1 : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                sel_idx[off + __popc(bal & ((1u << lane) - 1u))] = i;
                a.sum += (double)sv[u];
                a.n0 += (l == 0); a.n1 += (l == 1); a.n2 += (l == 2);
            }
            off += __popc(bal);
        }
    }
    count = total;
    tick(sm);   // emit: write pass done
    a = block_reduce_agg(a, sm);
    sel_sum = a.sum;
    sel_class[0] = a.n0; sel_class[1] = a.n1; sel_class[2] = a.n2;
}

__device__ __forceinline__ float position_bias_at(int i, int S, const float *tab, int tab_len, float log_S)
{
    if (S <= 1) return 0.0f;
    const float lg = (i < tab_len) ? tab[i] : (float)log((double)(float)(i + 1));
    return __fdiv_rn(lg, log_S);
}

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
    if (threadIdx.x == 0) { sm.dbg = p.dbg; sm.dbg_n = 0; tick(sm); }

    // ---- scores (token_importance.py:134-176) ----
    // Every loop below handles 4 independent tokens per trip (i, i+kT, i+2kT, i+3kT) with the
    // loads issued first: the per-token chain (L2 load -> 3 IEEE divisions -> stores) is long,
    // so instruction-level parallelism, not occupancy, is what keeps this single CTA moving.
    const float *__restrict__ ms = p.mass ? p.mass + (int64_t)b * S : nullptr;
    const float *__restrict__ tab = p.log_tab;
    float mn = 0.0f, den = 0.0f;
    if (p.flags & RTKV_FROM_MASS) {
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
                        float n = 0.0f;
                        if (ok) n = rnd_dt(__fdiv_rn(rnd_dt(__fsub_rn(s, mn), dt), den), dt);
                        const float t1 = rnd_dt(__fmul_rn(rnd_dt(__fmul_rn(p.alpha, n), dt), p.w_l), dt);
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
        if (all_bits <= cfg.budget_bits) {            // the whole row fits: no walk needed
            if (threadIdx.x == 0) { sm.thr[0] = sm.thr[1] = sm.thr[2] = ~0ull; sm.cur = all_bits; }
            __syncthreads();
        } else {
            select_thresholds(rsc, rlb, S, st.kmin, st.kmax, cfg, sm);
        }
        cost_bits = sm.cur;
        tick(sm);   // thresholds done
        emit_selection(rsc, rlb, S, p.mask + (int64_t)b * S, p.sel_idx + (int64_t)b * S, sm, count,
                       sel_class, sel_sum);
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

extern "C" int rtkv_score_select(const float *mass, float *scores, uint8_t *labels, uint8_t *mask,
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
    p.mass = mass; p.scores = scores; p.labels = labels; p.mask = mask; p.sel_idx = sel_idx;
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
    score_select_kernel<<<B, kT, smem, (cudaStream_t)stream>>>(p);
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
