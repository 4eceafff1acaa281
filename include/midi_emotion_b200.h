/* midi_emotion_b200.h -- C ABI of the B200-native hot path of serkansulun/midi-emotion.
 *
 * The reference has no FFI: its model path is Python calling stock PyTorch operators
 * (SURVEY.md 2.2).  Each entry point below replaces one group of those operator call sites;
 * the reference lines are cited per function (paths relative to /root/reference/src).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all calls only enqueue work (async);
 *   - return value 0 = ok, non-zero = error; text via me_last_error(); nothing throws;
 *   - the library owns no persistent device memory: callers (PyTorch) allocate everything;
 *   - "T" tensors are ME_F32 or ME_BF16 as selected by `dtype`; biases, LayerNorm affine
 *     parameters, statistics and the residual stream are always fp32;
 *   - matrices are row-major.  There is no CPU fallback: shape/alignment violations are errors.
 */
#ifndef MIDI_EMOTION_B200_H
#define MIDI_EMOTION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ME_F32 0
#define ME_BF16 1

/* conditioning modes (config.py:7-9) */
#define ME_COND_NONE 0
#define ME_COND_DISCRETE_TOKEN 1
#define ME_COND_CONTINUOUS_TOKEN 2
#define ME_COND_CONTINUOUS_CONCAT 3

/* attention implementations */
#define ME_ATTN_SIMT 0    /* exact-order fp32 SIMT kernels (fp32 parity path, any dtype) */
#define ME_ATTN_TENSOR 1  /* bf16 tensor-core kernels */

/* attention flags (me_attn_args.flags, me_layer_args.attn_flags) */
#define ME_ATTN_NONCAUSAL 1     /* every key is visible to every query (models/music_regression.py:78, no_mask=True:
                                 * mask = None); the relative term keeps its lower-triangular support, Srel[i,j] = 0
                                 * for j > i (_qe_masking, music_regression.py:256-262)                              */
#define ME_ATTN_REF_ROUNDING 2  /* ME_ATTN_TENSOR forward only: round QK^T, Srel, their sum and the scaled logits to
                                 * bf16 exactly where the reference does under autocast (music_multi.py:215-222);
                                 * default keeps them in fp32 (more accurate, fewer instructions)                    */

/* GEMM epilogue flags */
#define ME_EPI_BIAS 1        /* += bias[n] (fp32)                          */
#define ME_EPI_RELU 2        /* max(.,0)                                   */
#define ME_EPI_ADD_F32 4     /* += addend[m, n] (fp32, ld = ldd)           */
#define ME_EPI_RELU_MASK 8   /* zero where mask[m, n] <= 0 (T, ld = ldmask) */

const char* me_last_error(void);
int me_version(void);
/* number of CUDA kernels this library has launched in this process (benchmark bookkeeping) */
unsigned long long me_launch_count(void);
/* 1 when the running device is sm_100 (B200); the tcgen05 paths refuse to run otherwise. */
int me_device_is_sm100(void);
/* Live timing of the tcgen05 GEMM launches (bench.py's roofline figure): enable with a slot count,
 * run, then collect the summed CUDA-event durations, algorithmic FLOPs (2MNK) and launch count. */
int me_profile_enable(int capacity);
int me_profile_collect(double* total_ms, double* total_flops, int* launches);
/* The same per kernel class, without ending the collection: 0 = tcgen05 GEMMs, 1 = tensor-core attention forward,
 * 2 = attention backward key side (dK, dV), 3 = attention backward query side (dQ, dE).  Algorithmic FLOPs of the
 * attention classes: causal-minimum 2 (L^2 / 2) dh per product and head, three products each.
 * Call before me_profile_collect(), which resets. */
int me_profile_collect_class(int cls, double* total_ms, double* total_flops, int* launches);
/* sizeof() of the argument structs below, for binding self-checks (ctypes/cgo/JNI mirrors). */
int me_sizeof_attn_args(void);
int me_sizeof_attn_bwd_args(void);
int me_sizeof_layer_args(void);
int me_sizeof_layer_bwd_args(void);
int me_sizeof_decode_layer_args(void);

/* ---------------------------------------------------------------------------------------
 * Input stage.  Replaces models/music_multi.py:89-102 (generate_mask, Embedding, *sqrt(d-dc),
 * fc_condition, expand+cat, += positional table, dropout) and
 * models/music_continuous_token.py:81-100 (two Linear(1,d) prefix vectors, cat along sequence).
 *   tokens   int64 [B, L]            cond  f32 [B, 2] (never read when mode is none/discrete)
 *   emb_w    f32 [V, d - d_cond]     pe    f32 [max_seq, d]
 *   cw0/cb0  concat: fc_condition.weight [d_cond, 2] / bias [d_cond]
 *            ctoken: fc_condition.0.weight [d, 1] / bias [d];  cw1/cb1: fc_condition.1.*
 *   x_f32    out f32 [B, Ls, d]  (Ls = L + 2 for continuous_token, else L)
 *   x_T      out T   [B, Ls, d]  (may be NULL, or alias x_f32 when dtype == ME_F32)
 *   keypad   out u8  [B, Ls]     1 where the key is a pad token (mask[b,q,k] = k>q || keypad[b,k])
 * ------------------------------------------------------------------------------------- */
int me_embed_forward(const int64_t* tokens, const float* cond, const float* emb_w, const float* cw0,
                     const float* cb0, const float* cw1, const float* cb1, const float* pe, int B, int L,
                     int d, int d_cond, int V, int mode, int pad_token, float dropout_p, uint64_t seed,
                     int dtype, float* x_f32, void* x_T, uint8_t* keypad, void* stream);

/* Gradient of the input stage w.r.t. embedding.weight and fc_condition.* (autograd of the
 * lines above; train.py:317).  d_emb, d_cw0/1 and d_cb0/1 are fp32 and must be zeroed by the
 * caller (they accumulate). */
int me_embed_backward(const float* dx, const int64_t* tokens, const float* cond, int B, int L, int d,
                      int d_cond, int V, int mode, int pad_token, float dropout_p, uint64_t seed,
                      float* d_emb, float* d_cw0, float* d_cb0, float* d_cw1, float* d_cb1, void* stream);
/* The same with the gradient given as dx + dx_T (dx_T: bf16 [B, Ls, d] or NULL) -- the two-part gradient
 * me_layer_backward leaves (me_layer_bwd_args.d_x_T); saves a pass over [M, d] that would add them first. */
int me_embed_backward_split(const float* dx, const void* dx_T, const int64_t* tokens, const float* cond, int B, int L,
                            int d, int d_cond, int V, int mode, int pad_token, float dropout_p, uint64_t seed,
                            float* d_emb, float* d_cw0, float* d_cb0, float* d_cw1, float* d_cb1, void* stream);

/* ---------------------------------------------------------------------------------------
 * Linear layers.  D[M,N] = A . B^T (+ epilogue).  Replaces every torch.nn.Linear on the path
 * (music_multi.py:196-206,237,131-132,106) and their autograd (dgrad / wgrad).
 *   a_mn = 0: A is [M, K] with K contiguous (lda = row pitch);  a_mn = 1: A is stored [K, M].
 *   b_mn = 0: B is [N, K] with K contiguous;                    b_mn = 1: B is stored [K, N].
 *   out_dtype selects D's type.  bf16: tcgen05 + TMA (sm_100a).  f32: exact-order SIMT.
 * ------------------------------------------------------------------------------------- */
int me_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                 int a_mn, int b_mn, int out_dtype, int epi_flags, const float* bias, const float* addend,
                 const void* relu_mask, int ldmask, void* stream);
/* tuning/test hook: same as me_gemm_bf16 with a forced tile width (32/64/128/256, 0 = auto) and
 * split-K factor (0 = auto). */
int me_gemm_bf16_ex(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                    int a_mn, int b_mn, int out_dtype, int epi_flags, const float* bias, const float* addend,
                    const void* relu_mask, int ldmask, int tile_n, int splits, void* stream);
/* test hook: fp32 D = bf16 A . bf16 B^T on CUDA cores (device-side reference for the tcgen05 GEMM). */
int me_gemm_bf16_reference(const void* A, const void* B, float* D, int M, int N, int K, int lda, int ldb,
                           int ldd, int a_mn, int b_mn, void* stream);
int me_gemm_f32(const float* A, const float* B, float* D, int M, int N, int K, int lda, int ldb, int ldd,
                int a_mn, int b_mn, int epi_flags, const float* bias, const float* addend,
                const float* relu_mask, int ldmask, void* stream);

/* ---------------------------------------------------------------------------------------
 * out = LayerNorm(x_res + dropout(y)) ; post-LN residual block tail (music_multi.py:128-129,
 * 133-134; eps 1e-6).  y is T, everything else fp32.  z/mean/rstd are saved for backward when
 * non-NULL.  out_T may be NULL (or alias out_f32 for ME_F32).
 * ------------------------------------------------------------------------------------- */
int me_add_layernorm_forward(const float* x_res, const void* y, int dtype, const float* gamma,
                             const float* beta, float eps, int M, int d, float dropout_p, uint64_t seed,
                             float* out_f32, void* out_T, float* z, float* mean, float* rstd, void* stream);
/* dz = dLN(dout); d_gamma/d_beta accumulate (+=, caller zeroes).  dz_f32 feeds the residual
 * branch, dy_T = dropout_mask * dz feeds the sub-layer GEMMs.  dout_add (optional) is added to
 * dout first (the residual gradient that bypassed the following sub-layer). */
int me_add_layernorm_backward(const float* dout, const float* dout_add, const float* z, const float* mean,
                              const float* rstd, const float* gamma, int M, int d, float dropout_p,
                              uint64_t seed, int dtype, float* dz_f32, void* dy_T, float* d_gamma,
                              float* d_beta, void* stream);

/* column sums (bias gradients): out[n] += sum_m X[m, n]; X is T with row pitch ldx. */
int me_colsum(const void* X, int dtype, int M, int N, int ldx, float* out, void* stream);
/* the same with caller scratch (>= 148 * N floats, 16-byte aligned): partial rows + a finish pass instead of
 * same-address atomics, which serialise in the L2; falls back to me_colsum when the shape does not fit. */
int me_colsum_ws(const void* X, int dtype, int M, int N, int ldx, float* out, float* ws, int64_t ws_floats,
                 void* stream);
/* dst (T_dst, pitch ld_dst) = src (T_src, pitch ld_src), [rows, cols]; pad columns are zeroed. */
int me_convert_2d(const void* src, int src_dtype, int ld_src, void* dst, int dst_dtype, int ld_dst, int rows,
                  int cols, void* stream);

/* The same for a whole table of copies in ONE launch (the refresh of the compute-type weight copies at
 * every forward pass: the reference re-casts its fp32 parameters under autocast on every call too,
 * train.py:281).  `table_dev` is a DEVICE array of n entries, owned by the caller. */
typedef struct me_convert_desc {
  const void* src;
  void* dst;
  int32_t rows, cols, ld_src, ld_dst, src_dtype, dst_dtype;
} me_convert_desc;
int me_convert_batched(const me_convert_desc* table_dev, int n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Relative global attention (Music Transformer), causal + key-pad mask, fused:
 *   S[i,j] = (q_i.k_j + q_i.E[max_seq-1-(i-j)]) / sqrt(dh),  j <= i and !keypad[b,j]
 *   O = softmax(S) V
 * With ME_ATTN_NONCAUSAL every key j < Lk is visible (regression side model) and the E term is 0 for j > i.
 * Replaces music_multi.py:211-235 (_get_left_embedding, einsum, _qe_masking, _skewing, QK^T,
 * mask add, softmax, .V, head merge).  q/k/v are addressed with element strides so that they
 * may live in a packed QKV buffer [B, Lq, 3, H, dh] or in a KV cache [B, H, T, dh].
 *   q_pos0: absolute position of query row 0 (0 for full sequences, t for a decode step)
 *   Lq query rows per sequence, Lk keys per sequence (Lk = q_pos0 + Lq for self-attention)
 *   out [B, Lq, H*dh] T (pitch out_ld per row), lse f32 [B, H, Lq] (may be NULL when impl=SIMT)
 * ------------------------------------------------------------------------------------- */
typedef struct me_attn_args {
  int32_t dtype, impl;
  int32_t B, H, Lq, Lk, dh, max_seq, q_pos0, flags; /* flags: ME_ATTN_* bits */
  const void *q, *k, *v, *E;
  int64_t q_sb, q_sh, q_si; /* element strides: batch, head, row */
  int64_t k_sb, k_sh, k_sj;
  int64_t v_sb, v_sh, v_sj;
  const uint8_t* keypad; /* [B, keypad_ld] or NULL */
  int64_t keypad_ld;
  void* out;
  int64_t o_sb, o_si; /* out[b, i, h*dh + c] = out + b*o_sb + i*o_si + h*dh + c */
  float* lse;
  /* optional device-resident position (CUDA-graph friendly decode): when non-NULL,
   * q_pos0 = *pos_dev and Lk = *pos_dev + Lq are read on the device. */
  const int32_t* pos_dev;
  void* stream;
  /* Optional, ME_ATTN_TENSOR only: the forward pass also leaves its (unnormalised, bf16) probability tiles and the
   * per-row exponent offsets they were formed with, and the backward pass reads them back instead of recomputing
   * QK^T, the relative band, the skew and the exponentials (the reference keeps the whole softmax output for
   * autograd, music_multi.py:231).  Sizes: me_attention_saved_tiles().  NULL = recompute in backward. */
  void* p_tiles;  /* T [B * H * tiles, 128, 64]: 128 query rows x 64 keys per tile, UMMA K-major swizzled rows */
  float* m_tiles; /* f32 [B * H * tiles, 128]                                                                 */
} me_attn_args;
int me_attention_forward(const me_attn_args* a);
/* tiles per (sequence, head) of p_tiles / m_tiles for a sequence length and ME_ATTN_* flags */
int64_t me_attention_saved_tiles(int L, int flags);

/* Backward of the above (full self-attention only, q_pos0 = 0, Lq = Lk).
 *   dout T same addressing as out; dq/dk/dv T with the q/k/v strides; dE f32 [max_seq, dh]
 *   accumulates (+=, caller zeroes); dsum f32 [B,H,Lq] is scratch; dq_acc is fp32 scratch needed by
 *   ME_ATTN_TENSOR only (dq and dE accumulated across key tiles by TMA reduce-add): 16-byte aligned,
 *   me_attention_backward_workspace_floats(B, H, Lq, dh, max_seq) floats. */
typedef struct me_attn_bwd_args {
  me_attn_args f;
  const void* dout;
  void *dq, *dk, *dv;
  float* dE;
  float* dsum;
  float* dq_acc;
} me_attn_bwd_args;
int me_attention_backward(const me_attn_bwd_args* a);
int64_t me_attention_backward_workspace_floats(int B, int H, int L, int dh, int max_seq);
/* Debugging hook (kernel tuning only): when device_buf is non-NULL (>= 3*16*16 int64), one CTA of the
 * tensor-core attention backward kernel records clock64() stamps of its phases there; NULL disables. */
int me_debug_trace_set(long long* device_buf);

/* ---------------------------------------------------------------------------------------
 * One encoder layer (music_multi.py:126-135): attention block + FFN block, post-LN.
 * Weights: Wqkv = [Wq;Wk;Wv] packed [3d, d] T, E T [max_seq, dh], Wo [d,d], W1 [di,d], W2 [d,di];
 * biases / LN affine fp32.  Saved tensors are consumed by me_layer_backward.
 * ------------------------------------------------------------------------------------- */
typedef struct me_layer_args {
  int32_t dtype, attn_impl, training, attn_flags; /* attn_flags: ME_ATTN_* bits */
  int32_t B, Ls, d, H, d_inner, max_seq;
  float dropout_p, ln_eps;
  uint64_t seed;
  /* inputs */
  const float* x_f32;
  const void* x_T;
  const uint8_t* keypad;
  /* weights */
  const void* Wqkv;
  const float* bqkv;
  const void* E;
  const void* Wo;
  const float* bo;
  const float *ln1_w, *ln1_b;
  const void* W1;
  const float* b1;
  const void* W2;
  const float* b2;
  const float *ln2_w, *ln2_b;
  /* activations: saved (training) or scratch */
  void* qkv;      /* T [M, 3d] */
  void* attn_o;   /* T [M, d]  */
  float* lse;     /* [B, H, Ls] */
  void* proj;     /* T [M, d] scratch (attention out-proj, later FFN_suf output) */
  float* z1;      /* [M, d] pre-LN1 sum, NULL in eval */
  float *mean1, *rstd1;
  float* out1_f32;
  void* out1_T;
  void* h;        /* T [M, d_inner] */
  float* z2;
  float *mean2, *rstd2;
  float* out2_f32;
  void* out2_T;
  void* stream;
  /* optional saved attention probabilities (me_attn_args.p_tiles / m_tiles), ME_ATTN_TENSOR training only */
  void* attn_p;
  float* attn_m;
  /* Optional (ME_BF16 with saved statistics): the layer input handed over as the PREVIOUS LayerNorm's saved state
   * instead of its fp32 output: with xin_mean != NULL, x_f32 is that LayerNorm's pre-normalisation sum z and the
   * residual input is x = (z - xin_mean) * xin_rstd * xin_gamma + xin_beta, recomputed on the fly with the very
   * expression that LayerNorm evaluates -- the fp32 copy of a LayerNorm output (music_multi.py:129,134; read exactly
   * once, by the next LayerNorm) then never travels through HBM.  out1_f32 may be NULL when z1 / mean1 / rstd1 are
   * given (LN2 re-derives out1 the same way), out2_f32 may be NULL when the caller chains z2 / mean2 / rstd2 into
   * the next layer's xin_*.  All NULL: x_f32 is the input itself and both fp32 outputs are written. */
  const float *xin_mean, *xin_rstd, *xin_gamma, *xin_beta;
} me_layer_args;
int me_layer_forward(const me_layer_args* a);

typedef struct me_layer_bwd_args {
  me_layer_args f;      /* same tensors as the forward call */
  const float* d_out;   /* [M, d] fp32 gradient w.r.t. the layer output */
  float* d_x;           /* [M, d] fp32 gradient w.r.t. the layer input  */
  /* parameter gradients, fp32, written (not accumulated) except dE which must be zeroed */
  float *dWqkv, *dbqkv, *dE, *dWo, *dbo, *dln1_w, *dln1_b, *dW1, *db1, *dW2, *db2, *dln2_w, *dln2_b;
  /* scratch */
  float* g_a;           /* [M, d] fp32 */
  float* g_b;           /* [M, d] fp32 */
  void* g_T;            /* T [M, d]   */
  void* g_h;            /* T [M, d_inner] */
  void* g_qkv;          /* T [M, 3d]  */
  void* g_o;            /* T [M, d]   */
  float* dsum;          /* [B, H, Ls] */
  float* attn_ws;       /* me_attention_backward_workspace_floats(...) floats; ME_ATTN_TENSOR only */
  /* Optional (ME_BF16 only), the gradient as a sum of an fp32 and a compute-type part -- the rounding points of
   * the reference under autocast (a Linear's input gradient is bf16, the residual add is fp32):
   *   d_out_T  in : T [M, d] or NULL; the gradient w.r.t. the layer output is d_out + d_out_T
   *   d_x_T    out: T [M, d] or NULL; when given, the gradient w.r.t. the layer input is d_x + d_x_T
   *                 (d_x_T = the QKV projection's input gradient straight out of its GEMM) */
  const void* d_out_T;
  void* d_x_T;
} me_layer_bwd_args;
int me_layer_backward(const me_layer_bwd_args* a);

/* ---------------------------------------------------------------------------------------
 * KV-cache decode step for one layer (new capability; result must equal the reference's
 * full-prefix recompute, generate.py:99-122, while t < max_input_len <= max_seq).
 *   x_* [B, d] current-token activations; caches T [B, H, T_max, dh]; keypad [B, T_max].
 *   The step writes k/v of position t into the caches, attends over keys 0..t.
 * ------------------------------------------------------------------------------------- */
typedef struct me_decode_layer_args {
  me_layer_args f;  /* B, Ls = 1 ; x/out/qkv/attn_o/proj/h sized for M = B rows; keypad [B, T_max] */
  void* k_cache;
  void* v_cache;
  const int32_t* t_dev; /* device int32: position of the token being processed */
  int32_t T_max, _pad;
} me_decode_layer_args;
int me_decode_layer_step(const me_decode_layer_args* a);

/* Copy k/v rows of a packed QKV buffer [B, Ls, 3, H, dh] (prefill) into the caches at
 * positions pos0 .. pos0+Ls-1. */
int me_kv_cache_write(const void* qkv, int dtype, int B, int Ls, int H, int dh, void* k_cache, void* v_cache,
                      int T_max, int pos0, void* stream);

/* Decode-step input stage: x[b,:] for token tokens[b] at position *t_dev (same arithmetic as
 * me_embed_forward for a non-prefix position); also records keypad[b, *t_dev]. */
int me_embed_decode(const int64_t* tokens, const float* cond, const float* emb_w, const float* cw0,
                    const float* cb0, const float* pe, int B, int d, int d_cond, int V, int mode,
                    int pad_token, const int32_t* t_dev, int dtype, float* x_f32, void* x_T, uint8_t* keypad,
                    int T_max, void* stream);

/* ---------------------------------------------------------------------------------------
 * Next-token sampling of the generation loop, generate.py:122-189, as one launch with no host round
 * trips (the reference reads 2*B scalars back per generated token): NaN -> 0, special symbols -> -inf,
 * log_softmax, temperature (temperatures[0] when the token just fed is a TIMESHIFT tuple, else
 * temperatures[1]; plus the repeat penalty), top-k, top-p on the sorted cumulative softmax, softmax,
 * draw, repeat-count update.  The draw is inverse-CDF over the descending-sorted kept set with one
 * caller-supplied uniform in [0, 1) per sequence (torch.multinomial's generator stream is not
 * reproducible outside PyTorch).  V <= 4096.
 * ------------------------------------------------------------------------------------- */
typedef struct me_sample_args {
  int32_t B, V, ld_logits, logits_dtype; /* logits [B, ld_logits] T = ME_F32 | ME_BF16 (last position)     */
  const void* logits;
  const uint8_t* exclude;       /* [V] 1 = never sampled ("<...>" symbols, generate.py:57); may be NULL        */
  const uint8_t* is_timeshift;  /* [V] 1 = tuple token whose event is a TIMESHIFT (generate.py:143-147); NULL ok */
  const int64_t* prev_tokens;   /* [B] the token fed at this step (generate.py:140); NULL = rest temperature     */
  float temp_note, temp_rest;   /* temperatures[0], temperatures[1]                                             */
  float penalty_coeff;          /* generate.py:155-160; <= 0 disables                                           */
  int32_t top_k;                /* <= 0 or > V: all                                                             */
  float top_p;                  /* outside (0, 1): disabled                                                     */
  int32_t _pad;
  int32_t* repeat_counts;       /* [B] in/out (generate.py:185-189); may be NULL                                */
  const float* uniforms;        /* [B] in [0, 1)                                                                */
  int64_t* out_tokens;          /* [B]                                                                          */
  int32_t* out_num_choices;     /* [B] entries with non-zero probability; may be NULL                           */
  float* out_probs;             /* [B, V] final probabilities by token id; may be NULL (tests)                  */
  void* stream;
} me_sample_args;
int me_sample_step(const me_sample_args* a);
int me_sizeof_sample_args(void);

/* ---------------------------------------------------------------------------------------
 * Fused cross-entropy over the output head's logits: train.py:124,288-290 (CrossEntropyLoss with
 * ignore_index = pad, mean over the non-pad targets), its gradient, and the top-1 / top-5 hit counts of
 * utils.accuracy (utils.py:15-80, used at train.py:256), in one pass.
 *   logits T [M, ld] (V valid columns), targets int64 [M]
 *   grad_logits T [M, ld_grad] or NULL: d(mean loss)/d(logits); columns >= V are zeroed; may alias logits
 *   stats f32 [4] out: { sum of per-row losses, number of non-ignored rows, top-1 hits, top-5 hits }
 *   (mean loss = stats[0] / stats[1]; a target counts as a top-k hit when fewer than k logits are larger)
 * ------------------------------------------------------------------------------------- */
int me_cross_entropy(const void* logits, int dtype, int M, int V, int ld, const int64_t* targets,
                     int64_t ignore_index, void* grad_logits, int ld_grad, float* stats, void* stream);

/* ---------------------------------------------------------------------------------------
 * Output head fused with the training loss (bf16, tcgen05): models/music_multi.py:106 (self.fc) + train.py:288-290
 * (CrossEntropyLoss(ignore_index = pad)) + utils.py:15-80 (top-1 / top-5), without ever materialising the [M, V]
 * logits: the vocabulary is swept twice per 128-row tile (log-sum-exp, then gradient).
 *   x T=bf16 [M, K] (pitch ldx), W bf16 [V, K] (pitch ldw), bias f32 [V], targets int64 [M]
 *   grad_logits bf16 [M, ld_grad] or NULL: d(mean loss)/d(logits), columns >= V zeroed -- what the head's backward
 *               GEMMs read in place of the gradient me_cross_entropy would have written
 *   stats f32 [4] out: { sum of per-row losses, number of counted rows, top-1 hits, top-5 hits }
 * The logits are rounded to bf16 before the loss, exactly as the unfused path (and the reference under autocast)
 * holds them.
 * ------------------------------------------------------------------------------------- */
int me_head_cross_entropy(const void* x, const void* W, const float* bias, int M, int V, int K, int ldx, int ldw,
                          const int64_t* targets, int64_t ignore_index, void* grad_logits, int ld_grad, float* stats,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * Token pipeline of the training loader, data/loader.py:132-195 (Loader.__getitem__ after its random draws) with
 * data/data_processing.py:225-247 (transpose, tensor_to_ind_tensor), for a whole batch in one launch: transpose
 * the pitches of transposable events, map every (event, value) tuple to its token id, prepend the caller-built
 * prefix (<START> when the sample starts at a bar, <CLS>, the two emotion tokens of discrete_token), crop /
 * trim to input_len + 1, pad, and split into input = seq[:-1] and target = seq[1:] (left-padded by
 * target_left_pad pads: 2 for continuous_token, loader.py:189-193).  The random decisions (bar window,
 * n_transpose, crop offset) stay with the caller, so identical decisions give identical batches.
 * ------------------------------------------------------------------------------------- */
#define ME_TP_MAX_PREFIX 4
typedef struct me_token_pipeline_args {
  int32_t B, max_events;       /* samples; row pitch (in tuples) of `events`                                   */
  int32_t input_len;           /* loader.py:55-57: tgt_len, minus 2 for continuous_token                        */
  int32_t target_left_pad;     /* 0, or 2 for continuous_token                                                  */
  int32_t pad_token;
  int32_t n_event_types, n_values; /* dims of `lut`                                                             */
  int32_t min_pitch, max_pitch;    /* 21, 108 (data_processing.py:225)                                          */
  int32_t _pad;
  const int16_t* events;       /* [B, max_events, 2] (event index, value) tuples: the flattened bars            */
  const int32_t* n_events;     /* [B] tuples per sample                                                         */
  const int32_t* n_transpose;  /* [B] semitones, or NULL                                                        */
  const int32_t* start;        /* [B] crop offset >= 0, or -1 when the sample starts at a bar; NULL = all -1    */
  const int32_t* n_prefix;     /* [B] number of prefix tokens (<= ME_TP_MAX_PREFIX), or NULL                    */
  const int32_t* prefix;       /* [B, ME_TP_MAX_PREFIX] token ids in final order                                */
  const uint8_t* transposable; /* [n_event_types] 1 = pitched event (maps["transposable_event_inds"])           */
  const int32_t* lut;          /* [n_event_types, n_values] tuple -> token id (maps["tuple2idx"]), -1 = none    */
  int64_t* input;              /* out [B, input_len]                                                            */
  int64_t* target;             /* out [B, input_len + target_left_pad], or NULL (regression: loader.py:187-188) */
  int32_t* status;             /* out [B]: 1 when a tuple had no token id (the reference raises KeyError), or NULL */
  void* stream;
} me_token_pipeline_args;
int me_token_pipeline(const me_token_pipeline_args* a);
int me_sizeof_token_pipeline_args(void);

/* ---------------------------------------------------------------------------------------
 * Output head of the regression side model, models/music_regression.py:65-68,89 (selected by
 * models/build_model.py:29-32): out[b, :] = tanh(Linear(d, n_out)(x[b, 0, :])) -- the first position of every
 * sequence is pooled.  n_out <= 8 (the reference uses 2: valence, arousal).
 *   x  T [B, Ls, d] (fp32 for ME_F32; for ME_BF16 the bf16 copy of the last LayerNorm output, the weight is
 *      rounded to bf16 and the pre-activation / result are rounded to bf16 like the reference under autocast)
 *   W f32 [n_out, d], bias f32 [n_out], out f32 [B, n_out]
 * Backward: g_out f32 [B, n_out] -> dW [n_out, d], db [n_out] (written) and rows (b, 0) of d_x f32 [B, Ls, d]
 * (the other rows are not touched: the caller zeroes d_x).
 * ------------------------------------------------------------------------------------- */
int me_pooled_head_forward(const void* x, int dtype, const float* W, const float* bias, int B, int Ls, int d,
                           int n_out, float* out, void* stream);
int me_pooled_head_backward(const float* g_out, const float* out, const void* x, int dtype, const float* W, int B,
                            int Ls, int d, int n_out, float* dW, float* db, float* d_x, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimiser step of the training loop, train.py:319-325:
 *     scaler.unscale_(optimizer); clip_grad_norm_(model.parameters(), clip); scaler.step(optimizer)
 * with optimizer = torch.optim.Adam (train.py:182), over ALL parameter tensors in a handful of launches
 * (the reference's sequence is ~10 foreach launches per 50-tensor chunk plus two host synchronisations
 * inside GradScaler).  Three calls:
 *   1. me_grad_sqnorm_partials: one fp32 partial sum of (grad / grad_scale)^2 per 8192-element chunk;
 *   2. me_adam_prepare (one thread block): total_norm = sqrt(sum of the partials), the clip coefficient
 *      min(1, max_grad_norm / (total_norm + 1e-6)) (torch.nn.utils.clip_grad_norm_), the skip decision
 *      (non-finite norm: GradScaler.step would skip the update, train.py:323) and, when not skipped,
 *      step += 1 and the bias-correction scalars of that step -- no host round trip anywhere;
 *   3. me_adam_update: g = grad / grad_scale * coef (+ weight_decay * p);
 *      m = m + (1 - beta1)(g - m); v = beta2 v + (1 - beta2) g^2;
 *      p -= (lr / (1 - beta1^step)) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)      (torch.optim.Adam)
 *      Gradients are read-only (the clipped values are not written back; the reference zeroes them next,
 *      train.py:325).
 * `tensors` is a HOST array (pointers travel in kernel-parameter space, 64 tensors per launch); all
 * tensor pointers are DEVICE fp32, contiguous.  stats: DEVICE f32[8], written by me_adam_prepare:
 *   [0] total_norm  [1] clip coefficient  [2] 1 if the step is skipped else 0  [3] step count after this call
 *   [4] lr / (1 - beta1^step)  [5] sqrt(1 - beta2^step)  [6..7] reserved
 * step_dev: DEVICE f32[1], the number of updates applied so far (torch keeps `step` as a float32 tensor too).
 * ------------------------------------------------------------------------------------- */
typedef struct me_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} me_adam_tensor;
/* number of partial sums me_grad_sqnorm_partials writes for this tensor list (negative on bad input) */
int64_t me_grad_sqnorm_chunks(const me_adam_tensor* tensors, int n);
int me_grad_sqnorm_partials(const me_adam_tensor* tensors, int n, double grad_scale, float* partials,
                            int64_t partials_capacity, void* stream);
/* n_partials == 0 (partials may be NULL): no norm was taken, total_norm is reported as 0 and nothing is clipped;
 * max_grad_norm <= 0: no clipping (the non-finite check still applies when partials are given). */
int me_adam_prepare(const float* partials, int64_t n_partials, double max_grad_norm, double lr, double beta1,
                    double beta2, float* step_dev, float* stats, void* stream);
int me_adam_update(const me_adam_tensor* tensors, int n, double beta1, double beta2, double eps,
                   double weight_decay, double grad_scale, const float* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIDI_EMOTION_B200_H */
