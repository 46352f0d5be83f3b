// score_select_cluster.cu — K2 across a thread-block cluster.
//
// The single-CTA K2 (score_select.cu) is instruction-throughput bound on its one SM
// (~170 warp instructions per token: IEEE divisions, per-op rounding, statistics, keys).
// Here one batch row is handled by a CLUSTER of 8 CTAs on 8 SMs: every token-parallel pass
// (mass range, scores/labels/statistics, histogram, candidate collection, mask + compaction)
// runs on a 1/8 chunk per CTA, and the CTAs exchange a few small records through distributed
// shared memory:
//   all-gather   each CTA writes its 48-byte record into slot [rank] of every CTA, cluster.sync
//   histogram    local shared-memory histograms, then a reduce-scatter over DSMEM into rank 0
//   walk         rank 0 runs the bucket scan and the literal walk of the <=3 boundary buckets
//                (same routines as the single-CTA kernel) and broadcasts a command record;
//                boundary-bucket candidates are appended to rank 0's list with DSMEM atomics
// Semantics and results are bit-identical to score_select_kernel (same keys, same thresholds,
// statistics combined in rank order); the tests run both kernels against the oracle.
#include <cooperative_groups.h>

#include "score_select.cuh"

namespace cg = cooperative_groups;

namespace rtkv {

constexpr int kCS = 8;               // CTAs per cluster (portable maximum)

struct XRec {                        // exchanged record
    double sum, sumsq;
    unsigned kmin, kmax;
    int n0, n1, n2, nan, count, pad;
};

enum { CMD_DONE = 0, CMD_COLLECT = 1, CMD_HIST = 2 };
struct Cmd {
    int op, f, shift, nb;
    unsigned long long lo;
    int alive[3], pad;
    unsigned long long thr[3];
    long long cur;
};

struct ClSmem {
    Smem s;                          // control block; s.hist = this CTA's local histogram
    int ghist[3][kNB];               // cluster-wide histogram (valid on rank 0)
    XRec xch[2][kCS];
    Cmd cmd;
    unsigned long long bar[2];       // cluster barrier: two alternating mbarriers, 8 remote arrivals per use
};

__device__ __forceinline__ XRec to_xrec(const RowAgg &a, int count)
{
    XRec r;
    r.sum = a.sum; r.sumsq = a.sumsq; r.kmin = a.kmin; r.kmax = a.kmax;
    r.n0 = a.n0; r.n1 = a.n1; r.n2 = a.n2; r.nan = a.nan; r.count = count; r.pad = 0;
    return r;
}

__device__ __forceinline__ void ctick(Smem &sm)
{
    if (threadIdx.x == 0 && sm.dbg && sm.dbg_n < 62) sm.dbg[1 + sm.dbg_n++] = clock64();
}

// Cluster-wide barrier with release/acquire semantics at cluster scope, built on one mbarrier
// per CTA: after a CTA-local __syncthreads, 8 threads of every CTA arrive (remotely) on the 8
// CTAs' mbarriers and everybody waits on its own.  barrier.cluster (cg::cluster_group::sync) has
// all 8192 threads of the cluster arrive, which measured ~3000 cycles per use here.
// Consecutive uses alternate between two mbarriers so that a fast CTA's arrival for use n+1 can
// never be counted towards use n of a slower CTA's barrier.  `parity` counts the uses.
__device__ __forceinline__ void cl_barrier(ClSmem &cs, unsigned &parity)
{
    __syncthreads();
    const unsigned local = (unsigned)__cvta_generic_to_shared(&cs.bar[parity & 1u]);
    const unsigned ph = (parity >> 1) & 1u;
    if (threadIdx.x < kCS) {
        unsigned remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((unsigned)threadIdx.x));
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(local), "r"(ph) : "memory");
    ++parity;
}

// all-gather of one record per CTA: afterwards cs.xch[buf][0..7] (buf = returned index) holds the
// records of all ranks, in rank order, in every CTA's shared memory
__device__ __forceinline__ int cl_exchange(cg::cluster_group &cl, ClSmem &cs, int rank, int &phase, unsigned &parity,
                                           const XRec &mine)
{
    const int buf = phase & 1;
    if (threadIdx.x < kCS) *cl.map_shared_rank(&cs.xch[buf][rank], threadIdx.x) = mine;
    cl_barrier(cs, parity);
    ++phase;
    return buf;
}

// local histogram of the chunk for one level, then reduce-scatter into rank 0's ghist
__device__ void cl_histogram(cg::cluster_group &cl, ClSmem &cs, int rank, unsigned &parity, const float *s_sc,
                             const uint8_t *s_lb, int beg, int end, unsigned long long lo, int shift, int nb)
{
    Smem &sm = cs.s;
    for (int x = threadIdx.x; x < 3 * kNB; x += kT) (&sm.hist[0][0])[x] = 0;
    __syncthreads();
#pragma unroll 4
    for (int i = beg + threadIdx.x; i < end; i += kT) {
        const unsigned long long k = walk_key(s_sc[i - beg], i);
        if (k >= lo) {
            const unsigned long long bk = (k - lo) >> shift;
            if (bk < (unsigned long long)nb) atomicAdd(&sm.hist[s_lb[i - beg]][(int)bk], 1);
        }
    }
    cl_barrier(cs, parity);                              // all local histograms complete
    constexpr int kSlice = 3 * kNB / kCS;                // 768 bins per CTA
    int *root_g = cl.map_shared_rank(&cs.ghist[0][0], 0);
    for (int x = rank * kSlice + threadIdx.x; x < (rank + 1) * kSlice; x += kT) {
        int tot = 0;
#pragma unroll
        for (int r = 0; r < kCS; ++r) tot += cl.map_shared_rank(&sm.hist[0][0], r)[x];
        root_g[x] = tot;
    }
    cl_barrier(cs, parity);                              // rank 0's ghist complete
}

__global__ void __launch_bounds__(kT, 1) score_select_cluster_kernel(K2Params p)
{
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    ClSmem &cs = *reinterpret_cast<ClSmem *>(dyn_smem);
    Smem &sm = cs.s;
    const int rank = (int)cl.block_rank();
    const int b = blockIdx.x / kCS, S = p.S;
    const int C = ((S + kCS - 1) / kCS + 127) / 128 * 128;      // tokens per CTA (multiple of 128)
    const int beg = min(S, rank * C), end = min(S, beg + C);
    float *s_sc = reinterpret_cast<float *>(dyn_smem + ((sizeof(ClSmem) + 15) / 16) * 16);
    uint8_t *s_lb = reinterpret_cast<uint8_t *>(s_sc + C);
    float *__restrict__ sc = p.scores + (int64_t)b * S;
    uint8_t *__restrict__ lb = p.labels + (int64_t)b * S;
    const float *__restrict__ ms = p.mass ? p.mass + (int64_t)b * S : nullptr;
    const float *__restrict__ tab = p.log_tab;
    int phase = 0;
    unsigned parity = 0;
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) {
        sm.dbg = (rank == 0 && blockIdx.x == 0) ? p.dbg : nullptr; sm.dbg_n = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&cs.bar[0])), "r"(kCS));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&cs.bar[1])), "r"(kCS));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl.sync();       // every CTA's mbarrier exists before anybody arrives on it remotely
    ctick(sm);

    // ---- mass range over the whole row ----
    float mn = 0.0f, den = 0.0f;
    if ((p.flags & RTKV_FROM_MASS) && p.mass_range) {
        // the exchange kernel already reduced the row's mass range (rtkv_shard_stats_combine): no pass, no cluster step
        const volatile unsigned *mr = p.mass_range + 4 * b;
        const unsigned kmin = mr[0], kmax = mr[1], nan = mr[2];
        mn = score_of_key(kmax);
        float mx = score_of_key(kmin);
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        den = rnd_dt(__fsub_rn(mx, mn), p.score_dt);
    } else if (p.flags & RTKV_FROM_MASS) {
        RowAgg a = agg_identity();
        for (int i0 = beg + threadIdx.x; i0 < end; i0 += 4 * kT) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * kT; v[u] = i < end ? ms[i] : ms[i0]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a.nan |= (v[u] != v[u]);
                const unsigned k = score_key_desc(v[u]);
                a.kmin = k < a.kmin ? k : a.kmin; a.kmax = k > a.kmax ? k : a.kmax;
            }
        }
        ctick(sm);   // 1 mass loop
        a = block_reduce_agg(a, sm, false);
        ctick(sm);   // 2 block reduce
        const XRec *all = cs.xch[cl_exchange(cl, cs, rank, phase, parity, to_xrec(a, 0))];
        ctick(sm);   // 3 allgather
        unsigned kmin = 0xffffffffu, kmax = 0u;
        int nan = 0;
#pragma unroll
        for (int r = 0; r < kCS; ++r) {
            kmin = all[r].kmin < kmin ? all[r].kmin : kmin; kmax = all[r].kmax > kmax ? all[r].kmax : kmax;
            nan |= all[r].nan;
        }
        mn = score_of_key(kmax);
        float mx = score_of_key(kmin);
        if (nan) mn = mx = __int_as_float(0x7fc00000);
        den = rnd_dt(__fsub_rn(mx, mn), p.score_dt);
    }

    // ---- scores, labels, statistics of the chunk ----
    RowAgg st = agg_identity();
    {
        const int dt = p.score_dt;
        const bool ok = den > 1e-8f;
        const float t3 = __fmul_rn(p.gamma, p.rel);
        const bool from_mass = (p.flags & RTKV_FROM_MASS) != 0, make_labels = (p.flags & RTKV_MAKE_LABELS) != 0;
        for (int i0 = beg + threadIdx.x; i0 < end; i0 += 4 * kT) {
            float sv[4], lg[4];
            int lv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kT;
                const bool in = i < end;
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
                if (i < end) {
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
                    s_sc[i - beg] = s; s_lb[i - beg] = (uint8_t)l;
                    st.n0 += (l == 0); st.n1 += (l == 1); st.n2 += (l == 2);
                    st.sum += (double)s; st.sumsq += (double)s * (double)s;
                    st.nan |= (s != s);
                    const unsigned k = score_key_desc(s);
                    st.kmin = k < st.kmin ? k : st.kmin; st.kmax = k > st.kmax ? k : st.kmax;
                }
            }
        }
    }
    ctick(sm);       // 4 score loop
    st = block_reduce_agg(st, sm);
    ctick(sm);       // 5 block reduce
    const XRec *all = cs.xch[cl_exchange(cl, cs, rank, phase, parity, to_xrec(st, 0))];
    ctick(sm);       // 6 allgather
    RowAgg row = agg_identity();                         // whole-row statistics, combined in rank order
#pragma unroll
    for (int r = 0; r < kCS; ++r) {
        row.sum += all[r].sum; row.sumsq += all[r].sumsq;
        row.kmin = all[r].kmin < row.kmin ? all[r].kmin : row.kmin;
        row.kmax = all[r].kmax > row.kmax ? all[r].kmax : row.kmax;
        row.n0 += all[r].n0; row.n1 += all[r].n1; row.n2 += all[r].n2; row.nan |= all[r].nan;
    }
    float smin = score_of_key(row.kmax), smax = score_of_key(row.kmin);
    if (row.nan) smin = smax = __int_as_float(0x7fc00000);

    // ---- budgeted selection: thresholds ----
    long long cost_bits = 0;
    unsigned long long thr0 = ~0ull, thr1 = ~0ull, thr2 = ~0ull;
    SelCfg cfg;
    cfg.bits[0] = p.bits[0]; cfg.bits[1] = p.bits[1]; cfg.bits[2] = p.bits[2];
    cfg.budget_bits = p.budget_bits;
    if (p.flags & RTKV_SELECT) {
        const long long all_bits = (long long)cfg.bits[0] * row.n0 + (long long)cfg.bits[1] * row.n1 +
                                   (long long)cfg.bits[2] * row.n2;
        if (all_bits <= cfg.budget_bits) {
            cost_bits = all_bits;
        } else {
            // level 0 is known to every CTA; deeper levels arrive in commands from rank 0
            unsigned long long lo = (unsigned long long)row.kmin << 32;
            const unsigned long long span = (((unsigned long long)row.kmax << 32) | (unsigned)(S - 1)) - lo;
            int shift = span ? (64 - __clzll((long long)span)) - 11 : 0;
            if (shift < 0) shift = 0;
            int nb = (int)(span >> shift) + 1;
            if (rank == 0 && threadIdx.x == 0) {
                for (int c = 0; c < 3; ++c) { sm.alive[c] = 1; sm.thr[c] = ~0ull; }
                sm.cur = 0; sm.depth = 0;
                sm.lvl_lo[0] = lo; sm.lvl_shift[0] = shift; sm.lvl_nb[0] = nb; sm.lvl_resume[0] = 0;
            }
            cl_histogram(cl, cs, rank, parity, s_sc, s_lb, beg, end, lo, shift, nb);
            ctick(sm);   // level-0 histogram
            for (;;) {
                ctick(sm);   // round start
                if (rank == 0) {
                    // decide the next command from the cluster-wide histogram of the current level
                    __syncthreads();
                    int op = CMD_DONE, f = 0;
                    if (sm.alive[0] | sm.alive[1] | sm.alive[2]) {
                        const int d = sm.depth;
                        scan_find_on(sm.lvl_resume[d], cfg, sm, cs.ghist);
                        f = sm.found;
                        if (f == INT_MAX) {              // everything left in this level fits
                            __syncthreads();
                            if (threadIdx.x == 0) { sm.cur += sm.total; if (d > 0) sm.depth = d - 1; }
                            __syncthreads();
                            op = d == 0 ? CMD_DONE : CMD_HIST;           // back to the parent level
                        } else {
                            const int m_alive = (sm.alive[0] ? cs.ghist[0][f] : 0) + (sm.alive[1] ? cs.ghist[1][f] : 0) +
                                                (sm.alive[2] ? cs.ghist[2][f] : 0);
                            const int sh = sm.lvl_shift[d];
                            __syncthreads();
                            if (threadIdx.x == 0) { sm.cur += sm.sum_before; sm.lvl_resume[d] = f + 1; }
                            __syncthreads();
                            if (m_alive <= kMaxCand || sh == 0 || d + 1 >= kMaxDepth) {
                                op = CMD_COLLECT;
                                if (threadIdx.x == 0) sm.ncand = 0;
                            } else {                     // refine the boundary bucket
                                op = CMD_HIST;
                                if (threadIdx.x == 0) {
                                    const int ns = sh > 11 ? sh - 11 : 0;
                                    sm.lvl_lo[d + 1] = sm.lvl_lo[d] + ((unsigned long long)f << sh);
                                    sm.lvl_shift[d + 1] = ns; sm.lvl_nb[d + 1] = 1 << (sh - ns);
                                    sm.lvl_resume[d + 1] = 0; sm.depth = d + 1;
                                }
                            }
                            __syncthreads();
                        }
                    }
                    if (threadIdx.x < kCS) {             // broadcast the command
                        Cmd c;
                        const int d = sm.depth;
                        c.op = op; c.f = f; c.shift = sm.lvl_shift[d]; c.nb = sm.lvl_nb[d]; c.lo = sm.lvl_lo[d];
                        c.alive[0] = sm.alive[0]; c.alive[1] = sm.alive[1]; c.alive[2] = sm.alive[2]; c.pad = 0;
                        c.thr[0] = sm.thr[0]; c.thr[1] = sm.thr[1]; c.thr[2] = sm.thr[2]; c.cur = sm.cur;
                        *cl.map_shared_rank(&cs.cmd, threadIdx.x) = c;
                    }
                }
                ctick(sm);   // decision made
                cl_barrier(cs, parity);
                ctick(sm);   // command sync
                const Cmd c = cs.cmd;
                if (c.op == CMD_DONE) { thr0 = c.thr[0]; thr1 = c.thr[1]; thr2 = c.thr[2]; cost_bits = c.cur; break; }
                if (c.op == CMD_COLLECT) {
                    // live tokens of boundary bucket c.f go to rank 0's candidate list
                    int *root_n = cl.map_shared_rank(&sm.ncand, 0);
                    unsigned long long *root_k = cl.map_shared_rank(&sm.cand_key[0], 0);
                    uint8_t *root_l = cl.map_shared_rank(&sm.cand_lab[0], 0);
#pragma unroll 4
                    for (int i = beg + threadIdx.x; i < end; i += kT) {
                        const unsigned long long k = walk_key(s_sc[i - beg], i);
                        if (k >= c.lo && ((k - c.lo) >> c.shift) == (unsigned long long)c.f) {
                            const int l = s_lb[i - beg];
                            if (c.alive[l]) {
                                const int pos = atomicAdd(root_n, 1);
                                if (pos < kMaxCand) { root_k[pos] = k; root_l[pos] = (uint8_t)l; }
                            }
                        }
                    }
                    cl_barrier(cs, parity);
                    if (rank == 0) process_candidates(sm.ncand < kMaxCand ? sm.ncand : kMaxCand, cfg, sm);
                } else {                                 // CMD_HIST: (re)build the histogram of level (lo, shift, nb)
                    cl_histogram(cl, cs, rank, parity, s_sc, s_lb, beg, end, c.lo, c.shift, c.nb);
                }
            }
        }
    }

    ctick(sm);   // thresholds known
    // ---- mask + compacted index list of the chunk ----
    int count_row = 0;
    RowAgg sel = agg_identity();
    bool all_fit = false;
    if (p.flags & RTKV_SELECT) {
        const long long all_bits = (long long)cfg.bits[0] * row.n0 + (long long)cfg.bits[1] * row.n1 +
                                   (long long)cfg.bits[2] * row.n2;
        all_fit = all_bits <= cfg.budget_bits;
    }
    if (all_fit) {                                       // every token kept, in order: no exchange needed
        uint8_t *mask = p.mask + (int64_t)b * S;
        int32_t *sel_idx = p.sel_idx + (int64_t)b * S;
        for (int i = beg + threadIdx.x; i < end; i += kT) { mask[i] = 1; sel_idx[i] = i; }
        count_row = S;
        sel.sum = row.sum; sel.n0 = row.n0; sel.n1 = row.n1; sel.n2 = row.n2;
    } else if (p.flags & RTKV_SELECT) {
        uint8_t *mask = p.mask + (int64_t)b * S;
        int32_t *sel_idx = p.sel_idx + (int64_t)b * S;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wchunk = ((C + 31) / 32 + 127) / 128 * 128;
        const int wbeg = min(end, beg + warp * wchunk), wend = min(end, wbeg + wchunk);
        int mine = 0;
        for (int base = wbeg; base < wend; base += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + 32 * u + lane;
                bool keep = false;
                if (i < wend) {
                    const int l = s_lb[i - beg];
                    keep = walk_key(s_sc[i - beg], i) < (l == 0 ? thr0 : (l == 1 ? thr1 : thr2));
                }
                mine += __popc(__ballot_sync(0xffffffffu, keep));
            }
        }
        int cta_total;
        int off = block_excl_scan(lane == 0 ? mine : 0, cta_total, sm);
        off = __shfl_sync(0xffffffffu, off, 0);
        all = cs.xch[cl_exchange(cl, cs, rank, phase, parity, to_xrec(agg_identity(), cta_total))];
        int base_off = 0;
#pragma unroll
        for (int r = 0; r < kCS; ++r) { if (r < rank) base_off += all[r].count; count_row += all[r].count; }
        off += base_off;
        for (int base = wbeg; base < wend; base += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + 32 * u + lane;
                bool keep = false;
                float s = 0.0f;
                int l = 0;
                if (i < wend) {
                    s = s_sc[i - beg]; l = s_lb[i - beg];
                    keep = walk_key(s, i) < (l == 0 ? thr0 : (l == 1 ? thr1 : thr2));
                    mask[i] = keep ? 1 : 0;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    sel_idx[off + __popc(bal & ((1u << lane) - 1u))] = i;
                    sel.sum += (double)s;
                    sel.n0 += (l == 0); sel.n1 += (l == 1); sel.n2 += (l == 2);
                }
                off += __popc(bal);
            }
        }
        sel = block_reduce_agg(sel, sm);
        all = cs.xch[cl_exchange(cl, cs, rank, phase, parity, to_xrec(sel, 0))];
        sel = agg_identity();
#pragma unroll
        for (int r = 0; r < kCS; ++r) { sel.sum += all[r].sum; sel.n0 += all[r].n0; sel.n1 += all[r].n1; sel.n2 += all[r].n2; }
    }
    // no CTA touches another CTA's shared memory after the last all-gather: non-root CTAs may leave
    if (rank != 0) return;
    if (p.mass_range && threadIdx.x == 0) {       // every CTA read the range before the cluster steps above: reset it for the next layer
        unsigned *mr = p.mass_range + 4 * b;
        mr[0] = 0xffffffffu; mr[1] = 0u; mr[2] = 0u;
    }
    ctick(sm);   // emit done
    if (threadIdx.x == 0 && sm.dbg) sm.dbg[0] = sm.dbg_n;

    if (threadIdx.x == 0) {
        rtkv_row_stats &r = p.row_stats[b];
        r.n_selected = count_row; r.n_selected_greedy = count_row;
        r.sel_class[0] = sel.n0; r.sel_class[1] = sel.n1; r.sel_class[2] = sel.n2;
        r.label_count[0] = row.n0; r.label_count[1] = row.n1; r.label_count[2] = row.n2;
        r.cost_bits = cost_bits; r.sel_score_sum = sel.sum;
        r.score_sum = row.sum; r.score_sumsq = row.sumsq; r.score_min = smin; r.score_max = smax;
    }

    // ---- last row's root CTA: layer statistics and the emergency fallback ----
    int mx = count_row;
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
        used_fb = 1;                                     // rare path: single CTA, rows from global memory
        __syncthreads();
        SelCfg fcfg;
        fcfg.bits[0] = fcfg.bits[1] = fcfg.bits[2] = 1;
        fcfg.budget_bits = p.fallback_topk;
        for (int r = 0; r < p.B; ++r) {
            const float *rs = p.scores + (int64_t)r * S;
            const uint8_t *rl = p.labels + (int64_t)r * S;
            RowAgg a = agg_identity();
            for (int i = threadIdx.x; i < S; i += kT) {
                const unsigned k = score_key_desc(rs[i]);
                a.kmin = k < a.kmin ? k : a.kmin; a.kmax = k > a.kmax ? k : a.kmax;
            }
            a = block_reduce_agg(a, sm, false);
            select_thresholds(rs, rl, S, a.kmin, a.kmax, fcfg, sm);
            int cnt, cls[3]; double ss;
            emit_selection(rs, rl, S, p.mask + (int64_t)r * S, p.sel_idx + (int64_t)r * S, sm, cnt, cls, ss);
            if (threadIdx.x == 0) p.row_stats[r].n_selected = cnt;
            mx = cnt > mx ? cnt : mx;
        }
    }
    if (threadIdx.x == 0) {
        p.layer_stats->max_selected = mx;
        p.layer_stats->used_fallback = used_fb;
        p.layer_stats->reserved[0] = 0;
        p.layer_stats->reserved[1] = 0;
    }
}

// Returns RTKV_EUNSUPPORTED when the cluster variant does not apply (caller falls back to the
// single-CTA kernel), RTKV_OK after a successful launch.
int launch_score_select_cluster(const K2Params &p, cudaStream_t st, bool force)
{
    const int S = p.S;
    // Measured on B200 (tools/bench_kernels.py): one CTA costs ~5 us + 1.3..1.8 ns/token, the cluster
    // ~11..21 us (a dozen ~2 us cross-CTA steps) + 0.1..0.2 ns/token: the cluster pays off from ~8k tokens.
    if (S < (force ? 1024 : 8192)) return RTKV_EUNSUPPORTED;
    const int C = ((S + kCS - 1) / kCS + 127) / 128 * 128;
    const size_t smem = ((sizeof(ClSmem) + 15) / 16) * 16 + 4 * (size_t)C + (size_t)C + 16;
    if (smem > 200 * 1024) return RTKV_EUNSUPPORTED;
    static size_t attr_bytes[64] = {0};
    int dev = 0;
    RTKV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_bytes[dev] < smem) {
        RTKV_CUDA(cudaFuncSetAttribute(score_select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_bytes[dev] = 200 * 1024;
    }
    RTKV_CUDA(launch_kernel(score_select_cluster_kernel, dim3((unsigned)(p.B * kCS)), dim3(kT), smem, st, kCS, p));
    return RTKV_OK;
}

} // namespace rtkv
