// score_select.cuh — device code shared by the single-CTA K2 kernel (score_select.cu) and the
// thread-block-cluster K2 kernel (score_select_cluster.cu): walk keys, row statistics,
// block reductions/scans, the bucketed greedy walk.  See score_select.cu for the algorithm.
#pragma once
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
    unsigned *mass_range;   // optional [B][4]: min / max of score_key_desc over the row's mass, NaN flag, pad - produced by the
                            // head-sharding exchange kernel; consumed and reset to (0xffffffff, 0, 0) here
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
static __device__ RowAgg block_reduce_agg(RowAgg a, Smem &sm, bool with_sums = true)
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
static __device__ int block_excl_scan(int v, int &total, Smem &sm)
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
static __device__ void build_hist(const float *sc, const uint8_t *lb, int S, Smem &sm)
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
static __device__ void scan_find_on(int b0, const SelCfg &cfg, Smem &sm, const int (*hist)[kNB])
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
        w[k] = (x >= b0 && x < nb) ? w0c * hist[0][x] + w1c * hist[1][x] + w2c * hist[2][x] : 0;
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

static __device__ void scan_find(int b0, const SelCfg &cfg, Smem &sm) { scan_find_on(b0, cfg, sm, sm.hist); }

// Sort the m collected candidates (sm.cand_*) by key and walk them literally
// (the reference's loop, selective_propagation.py:119-131), updating cur / alive / thr.
static __device__ void process_candidates(int m, const SelCfg &cfg, Smem &sm)
{
    if ((int)threadIdx.x < m) { // rank by counting: keys are unique
        const unsigned long long k = sm.cand_key[threadIdx.x];
        int rank = 0;
        for (int u = 0; u < m; ++u) rank += sm.cand_key[u] < k;
        sm.sort_key[rank] = k;
        sm.sort_lab[rank] = sm.cand_lab[threadIdx.x];
    }
    __syncthreads();
    // The literal walk without a sequential loop.  The walk only changes regime when a token does not fit, which retires
    // its class and every class at least as expensive, so that happens at most 3 times.  Each round: costs of the still
    // alive classes from position `start` on -> block prefix sum -> first position that overflows; everything before it is
    // taken, the classes are retired with that candidate's key as threshold, the next round resumes behind it.
    int start = 0;
    for (int round = 0; round < 4; ++round) {
        const int t = threadIdx.x;
        int c = 0, cost = 0;
        if (t < m && t >= start) { c = sm.sort_lab[t]; cost = sm.alive[c] ? cfg.bits[c] : 0; }
        if (threadIdx.x == 0) sm.found = INT_MAX;
        int total;
        const int excl = block_excl_scan(cost, total, sm);       // its barriers also order the sm.found reset
        const long long cur = sm.cur;
        if (cost > 0 && cur + excl + cost > cfg.budget_bits) atomicMin(&sm.found, t);
        __syncthreads();
        const int f = sm.found;
        if (f == INT_MAX) {                                       // everything left fits
            __syncthreads();
            if (threadIdx.x == 0) sm.cur = cur + total;
            break;
        }
        if (t == f) {
            sm.cur = cur + excl;
            for (int c2 = 0; c2 < 3; ++c2)
                if (sm.alive[c2] && cfg.bits[c2] >= cfg.bits[c]) { sm.alive[c2] = 0; sm.thr[c2] = sm.sort_key[f]; }
        }
        start = f + 1;
        __syncthreads();
    }
    __syncthreads();
}

// Materialise the live tokens of bucket bsel, sort them by key and walk them literally.
static __device__ void walk_bucket(const float *sc, const uint8_t *lb, int S, int bsel,
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
    process_candidates(sm.ncand < kMaxCand ? sm.ncand : kMaxCand, cfg, sm);
}

// Thresholds thr[0..2] of one row.  kmin/kmax: range of score_key_desc over the row.
static __device__ void select_thresholds(const float *sc, const uint8_t *lb, int S, unsigned kmin,
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
static __device__ void emit_selection(const float *sc, const uint8_t *lb, int S, uint8_t *mask,
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


// position bias log(i+1)/log(S) (token_importance.py:103-108); tab = (float)log((double)(i+1))
__device__ __forceinline__ float position_bias_at(int i, int S, const float *tab, int tab_len, float log_S)
{
    if (S <= 1) return 0.0f;
    const float lg = (i < tab_len) ? tab[i] : (float)log((double)(float)(i + 1));
    return __fdiv_rn(lg, log_S);
}

// score_select_cluster.cu: K2 over a cluster of 8 CTAs; RTKV_EUNSUPPORTED = use the single-CTA kernel
int launch_score_select_cluster(const K2Params &p, cudaStream_t st, bool force);

} // namespace rtkv
