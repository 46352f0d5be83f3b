/*
 * rtkv.h — C ABI of librtkv.so: the B200 (sm_100a) implementation of the
 * reference's per-layer prefill KV-compression hot path.
 *
 * Every entry point replaces one reference interface (paths relative to the
 * reference repository, EvelynHung-79/RealTime-KV-cache-Compression):
 *
 *   rtkv_prompt_mass            src/compression/token_importance.py:21-47
 *   rtkv_prompt_mass_qk         (no reference code: fused Q*K_prompt^T entry)
 *   rtkv_position_bias          src/compression/token_importance.py:87-110
 *   rtkv_normalize_mass         src/compression/token_importance.py:49-85
 *   rtkv_score_select           src/compression/token_importance.py:134-176
 *                               + dynamic_quantization.py:21-60
 *                               + selective_propagation.py:40-161,203-211
 *   rtkv_quant_gather           src/compression/dynamic_quantization.py:62-196
 *                               + selective_propagation.py:213-232
 *   rtkv_quant_params / rtkv_quantize_with_params
 *                               src/compression/dynamic_quantization.py:62-126
 *   rtkv_gather_scalars         src/compression/selective_propagation.py:231-232
 *   rtkv_quant_pack / rtkv_unpack_dequant
 *                               (no reference code: packed-code layout)
 *   rtkv_compress_layer_host    src/compression/unified_compressor.py:95-172 (host buffers)
 *
 * Conventions
 *   - Plain C types only.  Unless a parameter is marked [host], every pointer
 *     is a DEVICE pointer valid on the current CUDA device.
 *   - All work is enqueued on the caller's stream (cudaStream_t passed as
 *     void*); no entry point except rtkv_compress_layer_host synchronises.
 *   - Return value: 0 ok; <0 rtkv error (bad argument / unsupported); >0 a
 *     cudaError_t.  rtkv_last_error() gives a thread-local message.  The
 *     library never aborts the process and has no CPU fallback.
 *   - Inputs are borrowed and never written; outputs are caller-owned.
 */
#ifndef RTKV_H
#define RTKV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTKV_ABI_VERSION 1

#if defined(__GNUC__)
#define RTKV_API __attribute__((visibility("default")))
#else
#define RTKV_API
#endif

enum { RTKV_F32 = 0, RTKV_F16 = 1, RTKV_BF16 = 2 };

enum {
    RTKV_OK = 0,
    RTKV_EINVAL = -1,       /* bad argument */
    RTKV_EUNSUPPORTED = -2, /* shape / dtype / bit width outside the supported set */
    RTKV_ENOMEM = -3,
    RTKV_ENODEVICE = -4     /* no usable sm_100 device */
};

/* flags of rtkv_score_select */
enum {
    RTKV_FROM_MASS = 1,    /* scores are computed from `mass` (else `scores` is an input) */
    RTKV_MAKE_LABELS = 2,  /* labels are computed from the scores (else `labels` is an input) */
    RTKV_SELECT = 4,       /* run the budgeted selection */
    RTKV_FALLBACK = 8      /* apply the top-10% emergency fallback when no row selected anything */
};

/* Per-layer scalars (reference: configs/base_config.py:4-57).  bits = {low, medium, high}. */
typedef struct rtkv_layer_params {
    double alpha, beta, gamma; /* CompressionConfig.alpha/beta/gamma                        */
    double layer_weight;       /* config.layer_weights[layer_idx]                            */
    double theta_h, theta_m;   /* precision thresholds                                       */
    double budget;             /* seq_len * propagation_ratio (selective_propagation.py:96)  */
    int32_t bits[3];           /* low/medium/high_precision_bits                             */
    int32_t prompt_len;        /* P (unified_compressor.py:55)                               */
    int32_t score_dtype;       /* dtype of the attention tensor the mass came from: half
                                  dtypes round the attention term like torch does            */
    int32_t fallback_topk;     /* max(1, int(0.1*S)) (selective_propagation.py:207)          */
} rtkv_layer_params;

/* Written by rtkv_score_select, one per batch row. */
typedef struct rtkv_row_stats {
    int32_t n_selected;       /* tokens kept in this row (after fallback, if it fired)       */
    int32_t n_selected_greedy;/* tokens kept by the budgeted selection proper                */
    int32_t sel_class[3];     /* kept tokens per label (low, medium, high), greedy selection */
    int32_t label_count[3];   /* tokens per label in the row                                 */
    int64_t cost_bits;        /* 8 * sum of selected token costs                             */
    double sel_score_sum;     /* sum of the selected tokens' scores                          */
    double score_sum;         /* sum of all scores of the row                                */
    double score_sumsq;
    float score_min, score_max;
} rtkv_row_stats;

/* Written by rtkv_score_select, one per call. */
typedef struct rtkv_layer_stats {
    int32_t max_selected;     /* max over rows of n_selected = S' of the output tensors      */
    int32_t used_fallback;    /* 1 if the emergency top-k fallback replaced the selection    */
    int32_t reserved[2];
} rtkv_layer_stats;

RTKV_API int rtkv_abi_version(void);
RTKV_API const char *rtkv_last_error(void);
/* [host] out: name, SM count, compute capability of the current device. */
RTKV_API int rtkv_device_info(char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor);

/* Per-device scratch: a 256-byte header (whole-tensor min/max words) followed by the table
 * (float)log((double)(i+1)) of the position bias, shared by all layers and sequence lengths
 * up to S.  rtkv_scratch_init fills it once (enqueued on `stream`); a scratch sized for a
 * smaller S than a later call still works (entries past the table are computed in place). */
RTKV_API int rtkv_scratch_bytes(int B, int S, size_t *bytes);
RTKV_API int rtkv_scratch_init(void *scratch, size_t bytes, void *stream);

/* mass[b,i] = sum_p mean_h A[b,h,i,idx[p]]   (idx == NULL: idx[p] = p).
 * A: [B,H,S,*] with element strides sb, sh, si and unit stride in the last dim. */
RTKV_API int rtkv_prompt_mass(const void *A, int dtype, int B, int H, int S,
                     int64_t sb, int64_t sh, int64_t si,
                     const int64_t *idx, int P, float *mass, void *stream);

/* Head sharding (one rank holds H_local = n_groups*group_size consecutive heads; float32 only):
 * partials[b,i,g] = sum_p sum_{h in local group g} A[b,h,i,idx[p]],  [B,S,H_local/group_size].
 * The ranks' buffers are all-gathered into [R][B*S][8/R] (all_gather_into_tensor) and
 * rtkv_mass_from_partials (n = B*S, n_ranks = R, groups_per_rank = 8/R) adds the 8 groups in head
 * order: mass = (p0 + ... + p7) / H — bit-identical to rtkv_prompt_mass on the unsharded tensor.
 * (n_ranks = 1: a plain [n][n_groups] buffer.) */
RTKV_API int rtkv_prompt_mass_partials(const void *A, int dtype, int B, int H_local, int S,
                     int64_t sb, int64_t sh, int64_t si, const int64_t *idx, int P,
                     int group_size, float *partials, void *stream);
RTKV_API int rtkv_mass_from_partials(const float *partials, int64_t n, int n_ranks, int groups_per_rank, int H,
                     float *mass, void *stream);

/* Same quantity without a materialised attention matrix (tcgen05 tensor cores + TMA):
 * head_sums[b,h,i] = sum_{j<P, j<=i if causal} exp(scale * q[b,h,i,:].k[b,h,j,:] - lse[b,h,i])
 * Q: [B,H,S,D], Kp: [B,H,P,D] contiguous (bf16 or f16, D == 128, P <= 128), lse: [B,H,S] natural-log
 * row log-sum-exp of the layer's attention logits; head_sums: [B,H,S] f32 (head-major, like lse).
 * rtkv_mass_from_head_sums adds the heads in order and divides by H. */
RTKV_API int rtkv_prompt_mass_qk(const void *Q, const void *Kp, const float *lse, int dtype,
                        int B, int H, int S, int D, int P, float scale, int causal,
                        float *head_sums, void *stream);
RTKV_API int rtkv_mass_from_head_sums(const float *head_sums, int B, int H, int S, float *mass, void *stream);

RTKV_API int rtkv_position_bias(int S, float *pb, void *stream);
RTKV_API int rtkv_normalize_mass(const float *mass, int dtype, int B, int S, float *out, void *stream);

/* Scores -> labels -> budgeted selection for every row of a [B,S] layer (one launch).
 *   mass    [B,S] f32 in  (RTKV_FROM_MASS)            scores [B,S] f32 in/out
 *   labels  [B,S] u8 in/out (0 low, 1 medium, 2 high) mask   [B,S] u8 out (1 = kept)
 *   sel_idx [B,S] i32 out: kept token indices, ascending, first n_selected valid
 *   row_stats [B], layer_stats [1] (when B > 1, layer_stats->reserved must be zero on entry;
 *   it is left zero), scratch/scratch_bytes: an initialised scratch or NULL/0. */
RTKV_API int rtkv_score_select(const float *mass, float *scores, uint8_t *labels, uint8_t *mask,
                      int32_t *sel_idx, rtkv_row_stats *row_stats, rtkv_layer_stats *layer_stats,
                      const void *scratch, size_t scratch_bytes, int B, int S,
                      const rtkv_layer_params *params /*[host]*/, int flags, void *stream);

/* Per-token asymmetric min/max fake-quantisation of the kept rows, written compacted.
 *   X0, X1   [B,S,n] inputs (X1 may be NULL); Y0, Y1 [B,out_rows,n] outputs.
 *   sel_idx  [B,S] or NULL (identity: row r is token r, every token kept)
 *   row_stats: n_selected per row is read from it (NULL with sel_idx == NULL)
 *   labels   [B,S] u8 or NULL (NULL or bits == NULL: rows are copied unquantised)
 *   Output rows r >= n_selected are zero-filled when zero_pad != 0. */
RTKV_API int rtkv_quant_gather(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                      const uint8_t *labels, const int32_t *bits /*[host] [3] or NULL*/,
                      int out_rows, int zero_pad, void *Y0, void *Y1, void *stream);

/* Same, restricted to batch rows [b_begin, b_begin+b_count) and output rows
 * [row_begin, row_begin+row_count): lets a caller pipeline the gather with its copies. */
RTKV_API int rtkv_quant_gather_range(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                      const uint8_t *labels, const int32_t *bits /*[host] [3] or NULL*/,
                      int out_rows, int zero_pad, void *Y0, void *Y1,
                      int b_begin, int b_count, int row_begin, int row_count, void *stream);

/* KV-head sharding: the reference's scale spans ALL heads of a token row, so ranks first
 * compute the local min/max of their kept rows (minmax [B,out_rows,4] = min0,max0,min1,max1;
 * padding rows give +inf,-inf), reduce them over ranks (MIN / MAX), and quantise with the
 * reduced values.  negate_max stores -max instead of max, so ONE element-wise MIN all-reduce
 * combines the ranks (negation is exact).  With one rank this equals rtkv_quant_gather bit for bit. */
RTKV_API int rtkv_row_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats, int out_rows,
                      int negate_max, float *minmax, void *stream);
RTKV_API int rtkv_quant_gather_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                      const uint8_t *labels, const int32_t *bits /*[host]*/, int out_rows, int zero_pad,
                      const float *minmax, int negated_max, void *Y0, void *Y1, void *stream);

/* KV-head sharding, fused: min/max exchange inside the quantise kernel over peer-mapped memory
 * (NVLink).  mailbox_ptrs [host]: `world` device pointers, entry s = rank s's mailbox as mapped into
 * this process (symmetric memory / CUDA IPC), each rtkv_p2p_mailbox_bytes(B, out_rows, world) bytes,
 * zero-filled once; state: 4 zero-initialised uint32 in local device memory (epoch, ticket, error flag, reserved).
 * Every rank must call this with the same B, out_rows, selection and call sequence, launched concurrently (a CTA
 * waits for the peers' CTAs of the same row), and a collective (or barrier) must separate consecutive calls that
 * reuse a mailbox parity twice (the score all-gather does).  If a peer's words do not arrive within ~20 s the
 * kernel sets state[2] = 1 and continues with its local min/max (results of that call are then not the reference's);
 * it never traps.
 * Equals rtkv_quant_gather on the concatenated row bit for bit. */
RTKV_API int64_t rtkv_p2p_mailbox_bytes(int B, int out_rows, int world);
RTKV_API int rtkv_quant_gather_p2p(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                      const uint8_t *labels, const int32_t *bits /*[host]*/, int out_rows, int zero_pad,
                      const uint64_t *mailbox_ptrs /*[host]*/, int rank, int world, uint32_t *state,
                      void *Y0, void *Y1, void *stream);

/* KV-head sharding with ONE exchange per layer (reference couplings: token_importance.py:45 head mean
 * before min-max, dynamic_quantization.py:183-184 scale spans all heads).  Rank r of `world` (1, 2, 4 or 8) holds
 * H_local = (8/world)*group_size consecutive heads: A [B,H_local,S,>=P] float32 and K/V row slices [B,S,n_local].
 *   rtkv_shard_stats    streams A, K and V once and stores, for every token, the record
 *                       (kmin, kmax, vmin, vmax, prompt-mass partial of each local head group) into the exchange
 *                       buffer of EVERY rank (peer stores; xbuf_ptrs [host]: `world` device pointers, entry s = rank
 *                       s's buffer as mapped into this process, rtkv_shard_exchange_bytes(B,S,world) bytes each,
 *                       zero-filled once), then publishes an epoch flag to every rank.  No NCCL call is involved.
 *   rtkv_shard_combine  waits for the flags of all ranks and reduces the records in rank order:
 *                       mass [B,S] (bit-identical to rtkv_prompt_mass on the unsharded tensor) and
 *                       minmax [B,S,4] = (kmin, kmax, vmin, vmax) of the WHOLE row, indexed by token.
 *   rtkv_quant_gather_token_minmax   rtkv_quant_gather with those row min/max (equals rtkv_quant_gather on the
 *                       concatenated row bit for bit).
 * state: 4 zero-initialised uint32 in local device memory (epoch, ticket, error flag, reserved); state[2] != 0 after
 * a call means a peer's flag never arrived (the results of that call are undefined; nothing traps).  Every rank must
 * issue the same sequence of calls; consecutive calls alternate between two halves of the buffer, which is all the
 * inter-rank synchronisation the scheme needs (DESIGN.md "Multi-GPU"). */
RTKV_API int64_t rtkv_shard_exchange_bytes(int B, int S, int world);
RTKV_API int rtkv_shard_stats(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh, int64_t si,
                      int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                      const uint64_t *xbuf_ptrs /*[host]*/, int rank, int world, uint32_t *state, void *stream);
RTKV_API int rtkv_shard_combine(const void *xbuf_local, uint32_t *state, int B, int S, int H, int world,
                      float *mass, float *minmax, void *stream);
/* rtkv_shard_stats and rtkv_shard_combine in ONE kernel (real multi-GPU only: every rank must be running the same call
 * concurrently).  CTA c of every rank works on the same token blocks, so after pushing its records it raises a per-CTA
 * flag at every rank, waits for the flags of the peers' CTA c and reduces its own tokens: no grid-wide flag, no second
 * launch.  Needs the staged kernel (A a [B,H_local,S,P] slice, P % 4 == 0, P <= 128), else RTKV_EUNSUPPORTED.  Same
 * results as the two-call scheme; the two schemes may be mixed from call to call as long as all ranks agree. */
RTKV_API int rtkv_shard_stats_combine(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh, int64_t si,
                      int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                      const uint64_t *xbuf_ptrs /*[host]*/, int rank, int world, uint32_t *state, int H,
                      float *mass, float *minmax, void *stream);
/* Same, and additionally reduces the mass RANGE of every batch row into mass_range [B][4] u32 = (min, max of the selection
 * kernel's descending sort key over the row's mass, NaN flag, pad), which must hold (0xffffffff, 0, 0, 0) on entry.
 * rtkv_score_select_ranged consumes it (skipping its own pass over the mass and one cross-CTA step) and resets it. */
RTKV_API int rtkv_shard_stats_combine_ranged(const void *A, int a_dtype, int B, int H_local, int S, int64_t sb, int64_t sh,
                      int64_t si, int P, int group_size, const void *K, const void *V, int kv_dtype, int n_local,
                      const uint64_t *xbuf_ptrs /*[host]*/, int rank, int world, uint32_t *state, int H,
                      float *mass, float *minmax, uint32_t *mass_range, void *stream);
RTKV_API int rtkv_score_select_ranged(const float *mass, uint32_t *mass_range, float *scores, uint8_t *labels, uint8_t *mask,
                      int32_t *sel_idx, rtkv_row_stats *row_stats, rtkv_layer_stats *layer_stats,
                      const void *scratch, size_t scratch_bytes, int B, int S,
                      const rtkv_layer_params *params /*[host]*/, int flags, void *stream);
RTKV_API int rtkv_quant_gather_token_minmax(const void *X0, const void *X1, int dtype, int B, int S, int n,
                      const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                      const uint8_t *labels, const int32_t *bits /*[host]*/, int out_rows, int zero_pad,
                      const float *minmax_by_token, void *Y0, void *Y1, void *stream);

/* Prefetch [p, p+bytes) into L2 (one prefetch.global.L2 per 128-byte line).  Run it on a side stream
 * while rtkv_score_select (one SM, latency bound) leaves HBM idle, over the first K rows. */
RTKV_API int rtkv_prefetch_l2(const void *p, size_t bytes, void *stream);

/* Whole-tensor scale / zero-point (qparams[0] = scale, [1] = zero_point, in `dtype`,
 * stored as f32) and quantise-dequantise with given parameters. */
RTKV_API int rtkv_quant_params(const void *x, int dtype, int64_t numel, int bits, float *qparams,
                      void *scratch, void *stream);
RTKV_API int rtkv_quantize_with_params(const void *x, int dtype, int64_t numel, int bits,
                              const float *qparams, void *y, void *stream);

/* selected_scores [B,out_rows] f32 and selected_labels [B,out_rows] i64, zero padded. */
RTKV_API int rtkv_gather_scalars(const float *scores, const uint8_t *labels, const int32_t *sel_idx,
                        const rtkv_row_stats *row_stats, int B, int S, int out_rows,
                        float *sel_scores, int64_t *sel_labels, void *stream);

/* labels u8 -> i64 (the dtype the reference returns), and bool mask -> selection
 * lists, for callers that hand in their own mask. */
RTKV_API int rtkv_labels_to_i64(const uint8_t *labels, int64_t n, int64_t *out, void *stream);

/* Packed rows (layout: DESIGN.md "packed rows").  row_bytes = rtkv_packed_row_bytes(n, max bits).
 * codes [B,out_rows,row_bytes] u8, qparams [B,out_rows,2] f32 (scale, zero_point). */
RTKV_API int64_t rtkv_packed_row_bytes(int n, int bits);
RTKV_API int rtkv_quant_pack(const void *X, int dtype, int B, int S, int n,
                    const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                    const uint8_t *labels, const int32_t *bits /*[host]*/,
                    int out_rows, int64_t row_bytes, uint8_t *codes, float *qparams,
                    void *stream);
RTKV_API int rtkv_unpack_dequant(const uint8_t *codes, const float *qparams, int dtype, int B, int S, int n,
                        const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                        const uint8_t *labels, const int32_t *bits /*[host]*/,
                        int out_rows, int64_t row_bytes, void *Y, void *stream);

/* ---- extended scale granularities and outlier split ------------------------------------------
 * BASELINE.json north star, subsystem (2): "per-channel K and per-token V group-wise scale/zero
 * computation and outlier detection and split".  The reference has no such behaviour (its scale spans
 * the whole [H*D] row, dynamic_quantization.py:183-184 = RTKV_Q_TOKEN without outliers, which is what a
 * NULL or zeroed options struct selects); the semantics are defined by oracle/rtkv_oracle.c
 * orc_fake_quant_ex and summarised here:
 *   RTKV_Q_TOKEN        one (scale, zero-point) per kept row
 *   RTKV_Q_TOKEN_GROUP  one per kept row and group of group_size consecutive channels (divides n)
 *   RTKV_Q_CHANNEL      one (min, max) per channel and group of token_group (1..32) consecutive KEPT rows
 *                       of a batch row; every row derives scale/zero-point from it with its own bits
 *   outlier_tau_q8      0 = off, else 256*tau (tau >= 1), token modes only: an element is an outlier iff
 *                       |x| > tau * mean|x| of its group, the mean taken in integer fixed point relative to
 *                       the group's largest exponent (order independent, hence bit-reproducible).  Outliers
 *                       are excluded from the group's min/max and kept exactly.
 * mode[0] applies to X0 (keys), mode[1] to X1 (values). */
enum { RTKV_Q_TOKEN = 0, RTKV_Q_TOKEN_GROUP = 1, RTKV_Q_CHANNEL = 2 };
typedef struct rtkv_quant_options {
    int32_t mode[2];
    int32_t group_size;
    int32_t token_group;
    int32_t outlier_tau_q8;
    int32_t reserved[3];
} rtkv_quant_options; /* 32 bytes */

/* rtkv_quant_gather with options.  outlier_counts: optional [B,out_rows,2] i32 (X0, X1), number of
 * outliers kept exactly in each output row (0 for padding rows). */
RTKV_API int rtkv_quant_gather_ex(const void *X0, const void *X1, int dtype, int B, int S, int n,
                         const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                         const uint8_t *labels, const int32_t *bits /*[host]*/,
                         int out_rows, int zero_pad, const rtkv_quant_options *options /*[host]*/,
                         void *Y0, void *Y1, int32_t *outlier_counts, void *stream);

/* Packed rows with options, one tensor per call (tensor = 0: options->mode[0], 1: mode[1]).
 *   codes    [B,out_rows,row_bytes] as rtkv_quant_pack; split outliers hold code 0
 *   qparams  token modes: [B,out_rows,G,2] (scale, zero-point), G = rtkv_quant_groups();
 *            RTKV_Q_CHANNEL: [B,ceil(out_rows/token_group),n,2] (min, max)
 *   outliers (token modes with tau, else NULL/0): per output row the first outlier_cap outliers in channel
 *            order: out_idx [B,out_rows,cap] i32, out_val [B,out_rows,cap] f32 (slots past the count are zero), out_cnt [B,out_rows] i32;
 *            outliers past the cap are quantised like inliers (clamped to the code range). */
RTKV_API int rtkv_quant_groups(int n, const rtkv_quant_options *options, int tensor);
RTKV_API int rtkv_quant_pack_ex(const void *X, int dtype, int B, int S, int n,
                       const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                       const uint8_t *labels, const int32_t *bits /*[host]*/,
                       int out_rows, int64_t row_bytes, const rtkv_quant_options *options /*[host]*/, int tensor,
                       uint8_t *codes, float *qparams, int32_t *out_idx, float *out_val, int32_t *out_cnt,
                       int outlier_cap, void *stream);
RTKV_API int rtkv_unpack_dequant_ex(const uint8_t *codes, const float *qparams, const int32_t *out_idx,
                           const float *out_val, const int32_t *out_cnt, int outlier_cap,
                           int dtype, int B, int S, int n,
                           const int32_t *sel_idx, const rtkv_row_stats *row_stats,
                           const uint8_t *labels, const int32_t *bits /*[host]*/,
                           int out_rows, int64_t row_bytes, const rtkv_quant_options *options /*[host]*/, int tensor,
                           void *Y, void *stream);

/* Decode-step attention directly over packed rows (reference-mode packing: rtkv_quant_pack, one (scale, zero-point)
 * per kept row; SURVEY.md 8f rank 2/3 - the consumer of the compressed cache the reference returns at
 * src/models/modified_llama.py:165-166, without ever expanding the codes in memory):
 *   out[b,h,:] = sum_j softmax_j(sm_scale * q[b,h,:].K'[b,j,h/(H/Hkv),:]) * V'[b,j,h/(H/Hkv),:]
 * q [B,H,D] f32, D == 128; codes_k/codes_v [B,out_rows,row_bytes]; qparams [B,out_rows,2]; row_bits [B,out_rows] u8 = the
 * bit width each row was packed with; row_stats[b].n_selected rows of batch row b are valid.  `splits` CTAs share the
 * rows of one (batch row, head); workspace: rtkv_packed_attention_workspace_bytes(B,H,splits).  out [B,H,D] f32, lse
 * [B,H] f32 (natural log) or NULL. */
RTKV_API int64_t rtkv_packed_attention_workspace_bytes(int B, int H, int splits);
RTKV_API int rtkv_packed_decode_attention(const float *q, const uint8_t *codes_k, const float *qparams_k,
                                 const uint8_t *codes_v, const float *qparams_v, const uint8_t *row_bits,
                                 const rtkv_row_stats *row_stats, int64_t row_bytes, int B, int H, int Hkv, int D,
                                 int out_rows, float sm_scale, int splits, void *workspace, size_t workspace_bytes,
                                 float *out, float *lse, void *stream);

/* Prefill attention on tcgen05 tensor cores + TMA, for the two neighbours of the compression call:
 *   upstream (src/models/modified_llama.py:88-94): causal softmax(Q K^T * sm_scale) V without the [B,H,S,S] tensor, returning
 *     out, the natural-log row LSE [B,H,Sq] and head_sums[b,h,i] = sum_{j<P, visible} softmax_ij - the prompt-column mass the
 *     reference slices out of the materialised probabilities (token_importance.py:41-45); rtkv_mass_from_head_sums turns it
 *     into the mass rtkv_score_select consumes;
 *   downstream (src/models/modified_llama.py:125-142): the same kernel with Skv = S' kept tokens: with causal != 0 query i sees
 *     kept slots j <= i, which is the reference's "mask truncated to its first S' columns".
 * Q [B,H,Sq,128], K/V [B,Hkv,Skv,128] (H % Hkv == 0), bf16 or f16, any strides (q/k/v/o_strides [host]: batch, head, token
 * strides in ELEMENTS, multiples of 8; unit stride in the last dim) - e.g. the transposed view of a [B,S,H,128] projection.
 * out: same dtype, written with o_strides (so a [B,Sq,H,128] result needs no transpose).  lse / head_sums may be NULL.
 * Logits, softmax and accumulation in float32; P is rounded to the input dtype for the second product. */
RTKV_API int rtkv_flash_prefill(const void *Q, const void *K, const void *V, int dtype, int B, int H, int Hkv, int Sq, int Skv, int D,
                       const int64_t *q_strides, const int64_t *k_strides, const int64_t *v_strides,
                       float sm_scale, int causal, int P, void *out, const int64_t *o_strides,
                       float *lse, float *head_sums, void *stream);

/* Profiling aid: when set to a device buffer of 64 int64, every rtkv_score_select launch writes
 * buf[0] = number of ticks and buf[1..] = clock64() at its phase boundaries (row 0).  NULL: off. */
RTKV_API int rtkv_debug_k2_timing(void *device_buffer_64_i64);

/* Testing aid: which rtkv_score_select kernel runs.  0 = automatic (a cluster of 8 CTAs per batch
 * row from 8192 tokens, else one CTA), 1 = always one CTA, 2 = the cluster kernel whenever S >= 1024. */
RTKV_API int rtkv_debug_k2_mode(int mode);

/* Profiling aid for rtkv_prompt_mass_qk (results become meaningless): 0 normal, 1 no exp in the
 * epilogue, 2 no TMEM loads, 3 no MMAs. */
RTKV_API int rtkv_debug_qk_mode(int mode);

/* Tuning aid: kernel used for RTKV_Q_CHANNEL groups of 17..32 rows: 0 (default) re-reads the tile through L2, 1 stages it
 * in shared memory with cp.async.  Same results. */
RTKV_API int rtkv_debug_channel_variant(int variant);

/* Testing aid: rtkv_shard_stats kernel: 0 (default) the bulk-copy staged kernel when the shape allows, 1 always the
 * generic register kernel.  Same results. */
RTKV_API int rtkv_debug_shard_variant(int variant);
/* Testing aid: cap the persistent grid of the staged statistics kernel (0 = one CTA per SM), so that several emulated
 * ranks can be co-resident on one GPU for rtkv_shard_stats_combine. */
RTKV_API int rtkv_debug_shard_grid(int max_ctas);
/* Test harness: rtkv_shard_stats_combine for ALL `world` ranks in ONE cooperative launch on one GPU (blockIdx.y = rank;
 * kernels that wait for one another must not be separate launches on one device).  [host] arrays of `world` device
 * pointers; params_scratch: >= 4096 bytes of device memory.  Same kernel body as the per-rank entry. */
RTKV_API int rtkv_debug_shard_stats_combine_allranks(const void *const *A, int a_dtype, int B, int H_local, int S, int64_t sb,
                      int64_t sh, int64_t si, int P, int group_size, const void *const *K, const void *const *V, int kv_dtype,
                      int n_local, const uint64_t *xbuf_ptrs, int world, uint32_t *const *state, int H, float *const *mass,
                      float *const *minmax, uint32_t *const *mass_range /*NULL or [world]*/, void *params_scratch, void *stream);

/* Testing aid: bf16 rows of rtkv_quant_gather*: 1 (default) packed two-elements-per-instruction path for rows inside its
 * guard, 0 element-wise code only.  Same results (per device: call with the device current). */
RTKV_API int rtkv_debug_bf16_packed(int on);

/* Testing aid: which kernel rtkv_flash_prefill launches.  -1 (default): by shape -- 4 from 3072 keys on, else 0.
 * 0: one query tile per CTA, P in shared memory, both softmax warpgroups on every key tile (column halves), agreeing on the row
 *    max once per tile;
 * 1: same layout, independent online softmaxes per key half merged in the epilogue;
 * 2: two query tiles per CTA with P in tensor memory (A operand of the second product from TMEM);
 * 3: each softmax warpgroup owns every other key tile with its own accumulator, merged in the epilogue;
 * 4: alternating key tiles with the running max handed from tile to tile, ONE accumulator, P in tensor memory, three K/V stages,
 *    separate issuing threads for the two products.
 * Same results up to fp32 summation order. */
RTKV_API int rtkv_debug_flash_variant(int variant);
/* Tuning aid: share of the exponentials of an unmasked tile that the default rtkv_flash_prefill kernel computes with a
 * degree-4 polynomial on the FMA pipe instead of MUFU.EX2: every `every`-th pair (2, 3, 4) or none (0, the default: measured slower at every share). */
RTKV_API int rtkv_debug_flash_poly(int every);
/* Tuning aid: `counters` = device buffer of >= 40 int64 (or NULL = off).  While set, bf16 launches of the default kernel run an
 * instrumented build that leaves the cycles CTA 0 spent in each phase of its tile loop there (tools/profile_flash_phases.py). */
RTKV_API int rtkv_debug_flash_profile(void *counters);
/* Tuning aid: cycles by which the odd softmax warpgroup of the alternating-tile kernel (variant 4) starts its first exponentials
 * late, so that the two warps of a scheduler use the MUFU pipe in turn instead of together.  0 = no offset. */
RTKV_API int rtkv_debug_flash_skew(int cycles);
/* Tuning aid: cycles one thread needs to issue (counters[0]) and the tensor pipe to retire (counters[1]) `n` back-to-back
 * 128x128x16 tcgen05.mma of the shapes rtkv_flash_prefill uses, on `ctas` CTAs at once.  mode 0: A, B in shared memory, K-major;
 * 1: B MN-major; 2: A in tensor memory, B MN-major; 3: 8 of mode 0 then 8 of mode 2, repeated; 4: mode 0 with N = 256;
 * 5: mode 0 cycling through four accumulators.  counters: >= 2 int64 (device). */
RTKV_API int rtkv_debug_mma_rate(int mode, int n, int ctas, void *counters, void *stream);

/* Testing aid: launch the hot kernels with (1, default) or without (0) programmatic dependent launch. */
RTKV_API int rtkv_debug_pdl(int on);

/* One layer with HOST buffers (pinned or pageable): copies K, V and the first P
 * key columns of A to the device, runs the path, copies the results back and
 * synchronises.  K_out/V_out must hold B*S*n elements; they are filled as a
 * contiguous [B, *max_selected, n].  scores [B,S] f32, labels/mask [B,S] u8,
 * row_stats [B] may be NULL.  ctx: rtkv_ctx_create. */
typedef struct rtkv_ctx rtkv_ctx;
RTKV_API int rtkv_ctx_create(rtkv_ctx **ctx);
RTKV_API int rtkv_ctx_destroy(rtkv_ctx *ctx);
RTKV_API int rtkv_compress_layer_host(rtkv_ctx *ctx,
                             const void *K, const void *V, int dtype, int B, int S, int n,
                             const void *A, int a_dtype, int H,
                             int64_t sb, int64_t sh, int64_t si,
                             const rtkv_layer_params *params,
                             void *K_out, void *V_out, int32_t *max_selected,
                             float *scores, uint8_t *labels, uint8_t *mask,
                             rtkv_row_stats *row_stats, rtkv_layer_stats *layer_stats);

#ifdef __cplusplus
}
#endif
#endif /* RTKV_H */
