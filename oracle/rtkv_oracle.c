/*
 * rtkv_oracle.c — CPU restatement of the reference's per-layer prefill
 * KV-compression path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (realtime-kv-cache-compression_b200/) never links, imports or calls it.
 *
 * Every function is a plain sequential restatement of the reference's PyTorch
 * code (paths relative to /root/reference), written as scalar loops so that it
 * shares no structure with the CUDA kernels it checks:
 *
 *   orc_prompt_mass        src/compression/token_importance.py:21-47
 *   orc_importance_scores  src/compression/token_importance.py:49-176
 *   orc_assign_precision   src/compression/dynamic_quantization.py:21-60
 *   orc_fake_quant_rows    src/compression/dynamic_quantization.py:62-196
 *   orc_select_tokens      src/compression/selective_propagation.py:40-161
 *   orc_apply_selection    src/compression/selective_propagation.py:163-244
 *   orc_compress_layer     src/compression/unified_compressor.py:95-172
 *
 * Parity pinning: the reference ships no golden vectors, so
 * this restatement is pinned against the reference itself, executed in the
 * build container by oracle/pin_against_reference.py, which also writes the
 * fixtures under tests/golden/.  Two things the reference leaves to ATen are
 * *defined* here (and implemented identically by the kernels):
 *   (1) the order of the H*P-term prompt-mass reduction (see orc_prompt_mass_groups),
 *   (2) logf of the position bias = (float)log((double)x) (correctly rounded),
 *   (3) ties in the descending score sort break towards the lowest index.
 * Against the reference those two stages agree to a few ulp; every later
 * stage (labels, codes, selection, gather) is bit-exact given equal scores.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -fopenmp
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_F32 = 0, ORC_F16 = 1, ORC_BF16 = 2 };

/* ---------- element types: every op rounds to the tensor dtype ---------- */

static inline float bits_to_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline uint16_t f32_to_bf16_bits(float f) {
    uint32_t u = f_to_bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u); /* quiet NaN */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb; /* round to nearest even */
    return (uint16_t)(u >> 16);
}
static inline float bf16_bits_to_f32(uint16_t h) { return bits_to_f((uint32_t)h << 16); }

static inline uint16_t f32_to_f16_bits(float f) { _Float16 h = (_Float16)f; uint16_t b; memcpy(&b, &h, 2); return b; }
static inline float f16_bits_to_f32(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }

/* round a float to the nearest value representable in dtype (RNE) */
static inline float rnd(float x, int dt) {
    if (dt == ORC_BF16) return bf16_bits_to_f32(f32_to_bf16_bits(x));
    if (dt == ORC_F16) return f16_bits_to_f32(f32_to_f16_bits(x));
    return x;
}
static inline size_t esize(int dt) { return dt == ORC_F32 ? 4 : 2; }
static inline float ld(const void *p, int64_t i, int dt) {
    if (dt == ORC_F32) return ((const float *)p)[i];
    uint16_t b = ((const uint16_t *)p)[i];
    return dt == ORC_BF16 ? bf16_bits_to_f32(b) : f16_bits_to_f32(b);
}
static inline void st(void *p, int64_t i, float v, int dt) {
    if (dt == ORC_F32) ((float *)p)[i] = v;
    else ((uint16_t *)p)[i] = dt == ORC_BF16 ? f32_to_bf16_bits(v) : f32_to_f16_bits(v);
}

/* ---------- a5: prompt-attention mass  (token_importance.py:41-45) ----------
 * mass[b,i] = sum_p ( mean_h A[b,h,i,idx[p]] ).
 * The reference leaves the order of this H*P-term reduction to ATen; here it is
 * DEFINED (the kernels implement exactly this, bit for bit):
 *   heads are split into 8 groups of gs = ceil(H/8) consecutive heads (groups past
 *   H are empty and contribute +0.0).  Prompt position p belongs to "lane"
 *   (p mod 128) / 4.
 *   fp32:  per group g: per position the group's heads are added sequentially;
 *          each lane adds its positions in increasing p; the 32 lane sums are
 *          combined by an xor butterfly (16,8,4,2,1) -> partial[g].
 *          mass = (partial[0] + partial[1] + ... + partial[7]) / (float)H.
 *          (scalar per-group partials are what head-sharded ranks exchange, so
 *           1/2/4/8-GPU runs are bit-identical to the single-GPU result)
 *   half:  torch rounds the per-position head mean to the tensor dtype, so per
 *          position s = group sums added in group order (fp32), m = rnd(s/H);
 *          lanes add their m in increasing p; butterfly; final rnd.
 * idx == NULL means idx[p] = p.
 */
static float butterfly32(float *lane)
{
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof(float) * 32);
    }
    return lane[0];
}

/* partials: optional [B,S,8] output of the fp32 per-group partial sums */
void orc_prompt_mass_groups(const void *A, int adt, int B, int H, int S,
                            int64_t sb, int64_t sh, int64_t si,
                            const int64_t *idx, int P, float *mass, float *partials)
{
    const int gs = (H + 7) / 8;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < S; ++i) {
            float lane[32];
            const int64_t row = b * sb + (int64_t)i * si;
            if (adt == ORC_F32) {
                float tot = 0.0f;
                for (int g = 0; g < 8; ++g) {
                    for (int l = 0; l < 32; ++l) lane[l] = 0.0f;
                    const int h0 = g * gs, h1 = (h0 + gs < H) ? h0 + gs : H;
                    for (int p = 0; p < P; ++p) {
                        int64_t col = idx ? idx[p] : p;
                        float s = 0.0f;
                        for (int h = h0; h < h1; ++h) s = s + ld(A, row + h * sh + col, adt);
                        int l = (p & 127) >> 2;
                        lane[l] = lane[l] + s;
                    }
                    float part = butterfly32(lane);
                    if (partials) partials[((int64_t)b * S + i) * 8 + g] = part;
                    tot = g == 0 ? part : tot + part;
                }
                mass[(int64_t)b * S + i] = tot / (float)H;
            } else {
                for (int l = 0; l < 32; ++l) lane[l] = 0.0f;
                for (int p = 0; p < P; ++p) {
                    int64_t col = idx ? idx[p] : p;
                    float s = 0.0f;
                    for (int g = 0; g < 8; ++g) {
                        const int h0 = g * gs, h1 = (h0 + gs < H) ? h0 + gs : H;
                        float gsum = 0.0f;
                        for (int h = h0; h < h1; ++h) gsum = gsum + ld(A, row + h * sh + col, adt);
                        s = g == 0 ? gsum : s + gsum;
                    }
                    float m = rnd(s / (float)H, adt);
                    int l = (p & 127) >> 2;
                    lane[l] = lane[l] + m;
                }
                mass[(int64_t)b * S + i] = rnd(butterfly32(lane), adt);
            }
        }
    }
}

void orc_prompt_mass(const void *A, int adt, int B, int H, int S,
                     int64_t sb, int64_t sh, int64_t si,
                     const int64_t *idx, int P, float *mass)
{
    orc_prompt_mass_groups(A, adt, B, H, S, sb, sh, si, idx, P, mass, NULL);
}

/* head-sharded rank: per-group scalar partial sums of a local head slice (fp32),
 * partials [B,S,H_local/group_size]; same per-group order as orc_prompt_mass_groups */
void orc_prompt_mass_partials(const float *A, int B, int H_local, int S,
                              int64_t sb, int64_t sh, int64_t si, int P, int group_size, float *partials)
{
    const int ng = H_local / group_size;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i)
            for (int g = 0; g < ng; ++g) {
                float lane[32];
                for (int l = 0; l < 32; ++l) lane[l] = 0.0f;
                for (int p = 0; p < P; ++p) {
                    float s = 0.0f;
                    for (int h = g * group_size; h < (g + 1) * group_size; ++h)
                        s = s + A[b * sb + h * sh + (int64_t)i * si + p];
                    int l = (p & 127) >> 2;
                    lane[l] = lane[l] + s;
                }
                partials[((int64_t)b * S + i) * ng + g] = butterfly32(lane);
            }
}

/* head-sharded combine: partials [n, ng] (all groups, gathered) -> mass */
void orc_mass_from_partials(const float *partials, int64_t n, int ng, int H, float *mass)
{
    for (int64_t t = 0; t < n; ++t) {
        float tot = partials[t * ng];
        for (int g = 1; g < ng; ++g) tot = tot + partials[t * ng + g];
        mass[t] = tot / (float)H;
    }
}

/* Extension with no reference behaviour (parity unpinned by the reference): prompt mass from
 * Q, K_prompt and the row log-sum-exp.  Q [B,H,S,D], Kp [B,H,P,D] in dtype dt, lse [B,H,S].
 * partials[b,i,g] = sum_{h in g} sum_{j<P, j<=i} exp(scale*q.k - lse)   (double accumulation:
 * this is the tolerance reference of the tensor-core kernel, not a bit-exact contract). */
void orc_prompt_mass_qk(const void *Q, const void *Kp, const float *lse, int dt, int B, int H, int S, int D,
                        int P, double scale, int causal, int group_size, float *partials)
{
    const int ng = H / group_size;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i)
            for (int g = 0; g < ng; ++g) {
                double acc = 0.0;
                for (int h = g * group_size; h < (g + 1) * group_size; ++h) {
                    const int64_t qo = (((int64_t)b * H + h) * S + i) * D;
                    const double l = (double)lse[((int64_t)b * H + h) * S + i];
                    for (int j = 0; j < P && (!causal || j <= i); ++j) {
                        const int64_t ko = (((int64_t)b * H + h) * P + j) * D;
                        double dot = 0.0;
                        for (int d = 0; d < D; ++d) dot += (double)ld(Q, qo + d, dt) * (double)ld(Kp, ko + d, dt);
                        acc += exp(scale * dot - l);
                    }
                }
                partials[((int64_t)b * S + i) * ng + g] = (float)acc;
            }
}

/* correctly rounded fp32 log (definition (2) above) */
static inline float logf_cr(float x) { return (float)log((double)x); }

/* ---------- a4,a6,a7,a8: scores  (token_importance.py:49-176) ---------- */
void orc_position_bias(int S, float *pb) /* :87-110 */
{
    if (S <= 1) { for (int i = 0; i < S; ++i) pb[i] = 0.0f; return; }
    float div = (float)log((double)S);
    for (int i = 0; i < S; ++i) pb[i] = logf_cr((float)(i + 1)) / div;
}

float orc_context_relevance(int S, int P) /* :112-132 */
{
    double r = (double)P / (double)S;
    return (float)(r < 1.0 ? r : 1.0);
}

void orc_normalize(const float *mass, int adt, int B, int S, float *out) /* :49-85 */
{
    for (int b = 0; b < B; ++b) {
        const float *m = mass + (int64_t)b * S;
        float mn = m[0], mx = m[0];
        int has_nan = 0;
        for (int i = 0; i < S; ++i) {
            if (m[i] != m[i]) has_nan = 1;
            if (m[i] < mn) mn = m[i];
            if (m[i] > mx) mx = m[i];
        }
        if (has_nan) mn = mx = NAN;
        float den = rnd(mx - mn, adt);
        for (int i = 0; i < S; ++i) {
            float n = 0.0f;
            if (den > 1e-8f) n = rnd(rnd(m[i] - mn, adt) / den, adt);
            out[(int64_t)b * S + i] = n;
        }
    }
}

void orc_importance_scores(const float *mass, int adt, int B, int S, int P,
                           double alpha, double beta, double gamma, double w_l,
                           float *scores) /* :134-176 */
{
    float *pb = (float *)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1));
    float *nrm = (float *)malloc(sizeof(float) * (size_t)B * (size_t)(S > 0 ? S : 1));
    orc_position_bias(S, pb);
    orc_normalize(mass, adt, B, S, nrm);
    float r = orc_context_relevance(S, P);
    float t3 = (float)gamma * r;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i) {
            float t1 = rnd(rnd((float)alpha * nrm[(int64_t)b * S + i], adt) * (float)w_l, adt);
            float t2 = (float)beta * pb[i];
            scores[(int64_t)b * S + i] = (t1 + t2) + t3;
        }
    free(pb); free(nrm);
}

/* ---------- a9: labels  (dynamic_quantization.py:38-45) ---------- */
void orc_assign_precision(const float *scores, int64_t n, double theta_h, double theta_m,
                          uint8_t *labels, int64_t counts[3])
{
    float th = (float)theta_h, tm = (float)theta_m;
    counts[0] = counts[1] = counts[2] = 0;
    for (int64_t i = 0; i < n; ++i) {
        float s = scores[i];
        int high = s >= th;
        int med = (s >= tm) && (s < th);
        uint8_t l = high ? 2 : (med ? 1 : 0);
        labels[i] = l;
        counts[l]++;
    }
}

/* ---------- a10,a11: one row  (dynamic_quantization.py:62-126) ----------
 * qp[0]=scale, qp[1]=zero_point (both in the tensor dtype, held as float).
 */
static void quant_params_row(const void *x, int64_t off, int n, int dt, int bits, float qp[2])
{
    float mn = ld(x, off, dt), mx = mn;
    int has_nan = 0;
    for (int c = 0; c < n; ++c) {
        float v = ld(x, off + c, dt);
        if (v != v) has_nan = 1;
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    if (has_nan) mn = mx = NAN;
    if (mx == mn) { qp[0] = 1.0f; qp[1] = 0.0f; return; }             /* :83-86 */
    float qmax = (float)((1ll << bits) - 1);
    float scale = rnd(rnd(mx - mn, dt) / qmax, dt);                     /* :92 */
    float zp = rnd(0.0f - rnd(mn / scale, dt), dt);                     /* :93 (not rounded to int) */
    qp[0] = scale; qp[1] = zp;
}

static inline float quant_code(float v, float scale, float zp, float qmax_r, int dt)
{
    float q = rintf(rnd(rnd(v / scale, dt) + zp, dt));                  /* :120 half-to-even */
    if (q == q) { if (q < 0.0f) q = 0.0f; if (q > qmax_r) q = qmax_r; } /* :121 clamp keeps NaN */
    return q;
}

/* a12: per-token mixed precision fake-quant of rows [rows, n].  labels pick
 * bits[label] (bits = {low, medium, high}, :169-174).  Optional outputs:
 * codes (float, integer valued q of :120-121) and qparams [rows,2]. */
void orc_fake_quant_rows(const void *x, int dt, int64_t rows, int n,
                         const uint8_t *labels, const int bits[3],
                         void *out, float *codes, float *qparams)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int nb = bits[labels[r]];
        float qp[2];
        int64_t off = r * (int64_t)n;
        quant_params_row(x, off, n, dt, nb, qp);
        float qmax_r = rnd((float)((1ll << nb) - 1), dt);
        for (int c = 0; c < n; ++c) {
            float q = quant_code(ld(x, off + c, dt), qp[0], qp[1], qmax_r, dt);
            float dq = rnd(rnd(q - qp[1], dt) * qp[0], dt);             /* :124 */
            if (out) st(out, off + c, dq, dt);
            if (codes) codes[off + c] = q;
        }
        if (qparams) { qparams[2 * r] = qp[0]; qparams[2 * r + 1] = qp[1]; }
    }
}

/* ---------- a14,a15: greedy budgeted selection (selective_propagation.py:68-161) ---------- */
typedef struct { float s; int32_t i; } orc_key;

/* descending score, NaN first (torch sort order), ties -> lowest index (contract) */
static int key_before(const orc_key *a, const orc_key *b)
{
    int an = a->s != a->s, bn = b->s != b->s;
    if (an != bn) return an;
    if (!an && a->s != b->s) return a->s > b->s;
    return a->i < b->i;
}
static void merge_sort(orc_key *a, orc_key *tmp, int n)
{
    if (n < 2) return;
    int h = n / 2;
    merge_sort(a, tmp, h); merge_sort(a + h, tmp, n - h);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = key_before(&a[j], &a[i]) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(orc_key) * (size_t)n);
}

/* per-row outputs: counts[b], cost[b] (sum of selected costs, double),
 * sel_score_sum[b] (double sum of selected scores), class_counts[b*3+label]. */
void orc_select_tokens(const float *scores, const uint8_t *labels, int B, int S,
                       const int bits[3], double budget,
                       uint8_t *mask, int32_t *counts, double *cost,
                       double *sel_score_sum, int32_t *class_counts)
{
    orc_key *keys = (orc_key *)malloc(sizeof(orc_key) * (size_t)(S > 0 ? S : 1));
    orc_key *tmp = (orc_key *)malloc(sizeof(orc_key) * (size_t)(S > 0 ? S : 1));
    for (int b = 0; b < B; ++b) {
        const float *sc = scores + (int64_t)b * S;
        const uint8_t *lb = labels + (int64_t)b * S;
        uint8_t *mk = mask + (int64_t)b * S;
        for (int i = 0; i < S; ++i) { keys[i].s = sc[i]; keys[i].i = i; mk[i] = 0; }
        merge_sort(keys, tmp, S);                                       /* :113 */
        double cur = 0.0, ssum = 0.0;
        int32_t cnt = 0, cc[3] = {0, 0, 0};
        for (int k = 0; k < S; ++k) {                                   /* :119-131 */
            int i = keys[k].i;
            double c = (double)(float)((double)bits[lb[i]] / 8.0);      /* :54-64 float tensor -> .item() */
            if (cur + c <= budget) { mk[i] = 1; cur += c; cnt++; ssum += (double)sc[i]; cc[lb[i]]++; }
        }
        counts[b] = cnt;
        if (cost) cost[b] = cur;
        if (sel_score_sum) sel_score_sum[b] = ssum;
        if (class_counts) { class_counts[3 * b] = cc[0]; class_counts[3 * b + 1] = cc[1]; class_counts[3 * b + 2] = cc[2]; }
    }
    free(keys); free(tmp);
}

/* emergency fallback (:203-211): top-k by score in every row, ties -> lowest index */
void orc_topk_mask(const float *scores, int B, int S, int k, uint8_t *mask)
{
    orc_key *keys = (orc_key *)malloc(sizeof(orc_key) * (size_t)(S > 0 ? S : 1));
    orc_key *tmp = (orc_key *)malloc(sizeof(orc_key) * (size_t)(S > 0 ? S : 1));
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < S; ++i) { keys[i].s = scores[(int64_t)b * S + i]; keys[i].i = i; mask[(int64_t)b * S + i] = 0; }
        merge_sort(keys, tmp, S);
        for (int j = 0; j < k && j < S; ++j) mask[(int64_t)b * S + keys[j].i] = 1;
    }
    free(keys); free(tmp);
}

/* ---------- a16: gather in ascending index order, zero padded (:213-232) ---------- */
void orc_gather_rows(const void *x, int dt, int B, int S, int n, const uint8_t *mask,
                     int max_sel, void *out)
{
    size_t es = esize(dt);
    memset(out, 0, (size_t)B * (size_t)max_sel * (size_t)n * es);
    for (int b = 0; b < B; ++b) {
        int o = 0;
        for (int i = 0; i < S; ++i)
            if (mask[(int64_t)b * S + i]) {
                memcpy((char *)out + ((size_t)b * max_sel + o) * n * es,
                       (const char *)x + ((size_t)b * S + i) * n * es, (size_t)n * es);
                o++;
            }
    }
}

/* ---------- a1: one layer (unified_compressor.py:95-172) ----------
 * Literal order of the reference: score -> label -> fake-quant ALL tokens ->
 * select -> gather.  k_out/v_out must hold B*S*HD elements (worst case);
 * they are filled as [B, max_sel, HD].  Returns max_sel.
 * work_k/work_v: B*S*HD scratch for the fake-quantised tensors.
 */
int orc_compress_layer(const void *K, const void *V, int dt, int B, int S, int HD,
                       const void *A, int adt, int H, int64_t sb, int64_t sh, int64_t si, int P,
                       double alpha, double beta, double gamma, double w_l,
                       double theta_h, double theta_m, const int bits[3], double ratio,
                       float *scores, uint8_t *labels, uint8_t *mask, int32_t *counts,
                       void *work_k, void *work_v, void *k_out, void *v_out)
{
    int64_t n = (int64_t)B * S;
    float *mass = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int64_t lc[3];
    orc_prompt_mass(A, adt, B, H, S, sb, sh, si, NULL, P, mass);
    orc_importance_scores(mass, adt, B, S, P, alpha, beta, gamma, w_l, scores);
    orc_assign_precision(scores, n, theta_h, theta_m, labels, lc);
    orc_fake_quant_rows(K, dt, n, HD, labels, bits, work_k, NULL, NULL);
    orc_fake_quant_rows(V, dt, n, HD, labels, bits, work_v, NULL, NULL);
    double budget = (double)S * ratio;                                  /* selective_propagation.py:96 */
    orc_select_tokens(scores, labels, B, S, bits, budget, mask, counts, NULL, NULL, NULL);
    int max_sel = 0;
    for (int b = 0; b < B; ++b) if (counts[b] > max_sel) max_sel = counts[b];
    if (max_sel == 0) {                                                 /* :205-211 */
        int k = (int)((double)S * 0.1); if (k < 1) k = 1;
        orc_topk_mask(scores, B, S, k, mask);
        max_sel = k < S ? k : S;
        for (int b = 0; b < B; ++b) counts[b] = max_sel;
    }
    orc_gather_rows(work_k, dt, B, S, HD, mask, max_sel, k_out);
    orc_gather_rows(work_v, dt, B, S, HD, mask, max_sel, v_out);
    free(mass);
    return max_sel;
}

/* ---------- packed-code extension (no reference behaviour) ----------
 * Layout defined by this project (DESIGN.md "packed rows"): integer codes q of
 * orc_fake_quant_rows, little-endian within a byte, 8/bits codes per byte
 * (bits in {1,2,4,8}), or one little-endian uint16 per code for 9..16 bits.
 */
void orc_pack_codes(const float *codes, int64_t rows, int n, const uint8_t *labels,
                    const int bits[3], int64_t row_bytes, uint8_t *out)
{
    memset(out, 0, (size_t)rows * (size_t)row_bytes);
    for (int64_t r = 0; r < rows; ++r) {
        int nb = bits[labels[r]];
        uint8_t *o = out + r * row_bytes;
        for (int c = 0; c < n; ++c) {
            uint32_t q = (uint32_t)codes[r * (int64_t)n + c];
            if (nb > 8) { o[2 * c] = (uint8_t)(q & 0xff); o[2 * c + 1] = (uint8_t)(q >> 8); }
            else { int per = 8 / nb; o[c / per] |= (uint8_t)(q << ((c % per) * nb)); }
        }
    }
}

/* ---------- extended scale granularities + outlier split ----------
 * North-star extras with NO reference behaviour ("parity unpinned by the reference"):
 * this restatement DEFINES them; mode 0 without outliers is orc_fake_quant_rows (checked in tests).
 *   mode 0  one (scale, zero-point) per row                      (the reference, dynamic_quantization.py:183-184)
 *   mode 1  one per row and group of `group` consecutive channels (group divides n)
 *   mode 2  one (min, max) per channel and group of `tgroup` consecutive rows; each row of the group
 *           derives its own scale/zero-point from that pair and ITS bit width
 * Outlier split (modes 0/1, tau_q8 = 256*tau >= 256): an element is an outlier iff |x| > tau * mean|x|
 * of its group.  The mean is taken in integer fixed point so that it does not depend on the order of
 * the additions: with E the exponent field of the group's largest magnitude, q_i = |x_i| * 2^(30-E)
 * truncated to an integer (denormals -> 0), mean = floor(sum q_i / g), T = floor(tau_q8 * mean / 256),
 * outlier <=> q_i > T.  A group holding a NaN, or whose largest magnitude is denormal/zero, has no
 * outliers.  Outliers are kept exactly (dense output) and are left out of the group's min/max; in a
 * row only the first `cap` outliers in channel order are split (cap < 0: all) - later ones are
 * quantised like inliers, i.e. clamped to the code range.
 * Outputs (each optional): out [rows,n] tensor dtype; codes [rows,n] float (split outliers: 0);
 * qparams: modes 0/1 [rows, n/group, 2] = (scale, zero-point), mode 2 [ceil(rows/tgroup), n, 2] =
 * (min, max); omask [rows,n] = 1 for split outliers. */
static inline uint32_t mag_fix(uint32_t ab, uint32_t E)
{
    uint32_t e = ab >> 23;
    if (e == 0) return 0;
    uint32_t sh = E - e;
    if (sh >= 31) return 0;
    return (((ab & 0x7fffffu) | 0x800000u) << 7) >> sh;
}

static void qparams_from_minmax(float mn, float mx, int bits, int dt, float qp[2])
{
    if (mx == mn) { qp[0] = 1.0f; qp[1] = 0.0f; return; }
    float qmax = (float)((1ll << bits) - 1);
    qp[0] = rnd(rnd(mx - mn, dt) / qmax, dt);
    qp[1] = rnd(0.0f - rnd(mn / qp[0], dt), dt);
}

void orc_fake_quant_ex(const void *x, int dt, int64_t rows, int n, const uint8_t *labels, const int bits[3],
                       int mode, int group, int tgroup, int tau_q8, int cap,
                       void *out, float *codes, float *qparams, uint8_t *omask)
{
    if (mode == 2) {
        int64_t ng = (rows + tgroup - 1) / tgroup;
#pragma omp parallel for schedule(static)
        for (int64_t g = 0; g < ng; ++g) {
            int64_t r0 = g * tgroup, r1 = r0 + tgroup < rows ? r0 + tgroup : rows;
            for (int c = 0; c < n; ++c) {
                float mn = ld(x, r0 * n + c, dt), mx = mn;
                int has_nan = 0;
                for (int64_t r = r0; r < r1; ++r) {
                    float v = ld(x, r * n + c, dt);
                    if (v != v) has_nan = 1;
                    if (v < mn) mn = v;
                    if (v > mx) mx = v;
                }
                if (has_nan) mn = mx = NAN;
                if (qparams) { qparams[(g * n + c) * 2] = mn; qparams[(g * n + c) * 2 + 1] = mx; }
                for (int64_t r = r0; r < r1; ++r) {
                    int nb = bits[labels[r]];
                    float qp[2];
                    qparams_from_minmax(mn, mx, nb, dt, qp);
                    float qmax_r = rnd((float)((1ll << nb) - 1), dt);
                    float q = quant_code(ld(x, r * n + c, dt), qp[0], qp[1], qmax_r, dt);
                    if (out) st(out, r * n + c, rnd(rnd(q - qp[1], dt) * qp[0], dt), dt);
                    if (codes) codes[r * n + c] = q;
                    if (omask) omask[r * n + c] = 0;
                }
            }
        }
        return;
    }
    int g = mode == 1 ? group : n;
    int ngr = n / g;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int nb = bits[labels[r]];
        float qmax_r = rnd((float)((1ll << nb) - 1), dt);
        int64_t off = r * (int64_t)n;
        int split = 0;                                   /* outliers split so far in this row */
        for (int gi = 0; gi < ngr; ++gi) {
            int c0 = gi * g;
            int has_nan = 0;
            uint32_t maxab = 0;
            for (int c = c0; c < c0 + g; ++c) {
                float v = ld(x, off + c, dt);
                if (v != v) has_nan = 1;
                uint32_t ab = f_to_bits(v) & 0x7fffffffu;
                if (ab > maxab) maxab = ab;
            }
            uint64_t T = ~0ull;
            uint32_t E = maxab >> 23;
            int use_out = tau_q8 > 0 && !has_nan && E > 0;
            if (use_out) {
                uint64_t sum = 0;
                for (int c = c0; c < c0 + g; ++c) sum += mag_fix(f_to_bits(ld(x, off + c, dt)) & 0x7fffffffu, E);
                T = ((uint64_t)tau_q8 * (sum / (uint64_t)g)) >> 8;
            }
            float mn = INFINITY, mx = -INFINITY;
            for (int c = c0; c < c0 + g; ++c) {
                float v = ld(x, off + c, dt);
                if (use_out && mag_fix(f_to_bits(v) & 0x7fffffffu, E) > T) continue;
                if (v < mn) mn = v;
                if (v > mx) mx = v;
            }
            if (has_nan) mn = mx = NAN;
            float qp[2];
            qparams_from_minmax(mn, mx, nb, dt, qp);
            if (qparams) { qparams[(r * ngr + gi) * 2] = qp[0]; qparams[(r * ngr + gi) * 2 + 1] = qp[1]; }
            for (int c = c0; c < c0 + g; ++c) {
                float v = ld(x, off + c, dt);
                int is_out = use_out && mag_fix(f_to_bits(v) & 0x7fffffffu, E) > T;
                if (is_out && (cap < 0 || split < cap)) {
                    ++split;
                    if (out) st(out, off + c, v, dt);
                    if (codes) codes[off + c] = 0.0f;
                    if (omask) omask[off + c] = 1;
                    continue;
                }
                float q = quant_code(v, qp[0], qp[1], qmax_r, dt);
                if (out) st(out, off + c, rnd(rnd(q - qp[1], dt) * qp[0], dt), dt);
                if (codes) codes[off + c] = q;
                if (omask) omask[off + c] = 0;
            }
        }
    }
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
