// One encoder layer (music_multi.py:126-135) as a sequence of kernel launches on one stream:
// forward, backward (hand-derived, no autograd) and the KV-cache decode step.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

int launch_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                     int a_mn, int b_mn, int out_dtype, int flags, const float* bias, const float* addend,
                     const void* relu_mask, int ldmask, int force_bn, int force_splits, cudaStream_t st,
                     float* colsum_out, bool* colsum_done);
int launch_gemm_f32(const float* A, const float* B, float* D, int M, int N, int K, int lda, int ldb, int ldd,
                    int a_mn, int b_mn, int flags, const float* bias, const float* addend,
                    const float* relu_mask, int ldmask, cudaStream_t st);
int launch_add_ln_fwd(const float* x_res, const void* y, int dtype, const float* gamma, const float* beta,
                      float eps, int M, int d, float p, uint64_t seed, float* out_f32, void* out_T, float* z,
                      float* mean, float* rstd, const float* src_mean, const float* src_rstd, const float* src_gamma,
                      const float* src_beta, cudaStream_t st);
int launch_add_ln_bwd(const float* dout, const float* dout_add, const void* dout_add_T, const float* z,
                      const float* mean, const float* rstd, const float* gamma, int M, int d, float p, uint64_t seed,
                      int dtype, float* dz_f32, void* dy_T, float* d_gamma, float* d_beta, float* d_ybias,
                      cudaStream_t st);
int launch_colsum_ws(const void* X, int dtype, int M, int N, int ldx, float* out, float* ws, int64_t ws_floats,
                     cudaStream_t st);
int launch_kv_write(const void* qkv, int dtype, int B, int Ls, int H, int dh, void* kc, void* vc, int T_max,
                    int pos0, const int32_t* t_dev, cudaStream_t st);

// D = A . B^T in the layer's compute type.  out_f32 forces an fp32 result (residual-stream gradients).
static int linear(int dtype, const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                  int a_mn, int b_mn, bool out_f32, int flags, const float* bias, const float* addend,
                  const void* relu_mask, int ldmask, cudaStream_t st, float* colsum_out = nullptr,
                  bool* colsum_done = nullptr) {
  if (colsum_done) *colsum_done = false;
  if (dtype == ME_F32)
    return launch_gemm_f32(static_cast<const float*>(A), static_cast<const float*>(B), static_cast<float*>(D), M, N,
                           K, lda, ldb, ldd, a_mn, b_mn, flags, bias, addend,
                           static_cast<const float*>(relu_mask), ldmask, st);
  return launch_gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, a_mn, b_mn, out_f32 ? ME_F32 : ME_BF16, flags, bias,
                          addend, relu_mask, ldmask, 0, 0, st, colsum_out, colsum_done);
}

static inline size_t esize(int dtype) { return dtype == ME_BF16 ? 2 : 4; }
static inline const void* offs(const void* p, int dtype, int64_t elems) {
  return static_cast<const char*>(p) + elems * static_cast<int64_t>(esize(dtype));
}
static inline void* offs(void* p, int dtype, int64_t elems) {
  return static_cast<char*>(p) + elems * static_cast<int64_t>(esize(dtype));
}

static int check_layer(const me_layer_args* a, const char* who) {
  ME_CHECK(a != nullptr, "%s: NULL args", who);
  ME_CHECK(a->dtype == ME_F32 || a->dtype == ME_BF16, "%s: bad dtype %d", who, a->dtype);
  ME_CHECK(a->B > 0 && a->Ls > 0 && a->d > 0 && a->H > 0 && a->d_inner > 0, "%s: bad dims", who);
  ME_CHECK(a->d % a->H == 0, "%s: d_model %d not divisible by n_head %d", who, a->d, a->H);
  ME_CHECK(a->d % 8 == 0 && a->d_inner % 8 == 0, "%s: d_model and d_inner must be multiples of 8", who);
  ME_CHECK(a->x_f32 && a->x_T && a->Wqkv && a->bqkv && a->E && a->Wo && a->bo && a->W1 && a->b1 && a->W2 && a->b2 &&
               a->ln1_w && a->ln1_b && a->ln2_w && a->ln2_b,
           "%s: NULL input or weight pointer", who);
  ME_CHECK(a->qkv && a->attn_o && a->proj && a->out1_T && a->h && a->out2_T, "%s: NULL activation pointer", who);
  ME_CHECK(a->out1_f32 || (a->z1 && a->mean1 && a->rstd1), "%s: out1_f32 may only be NULL when z1/mean1/rstd1 are saved", who);
  ME_CHECK(a->out2_f32 || (a->z2 && a->mean2 && a->rstd2), "%s: out2_f32 may only be NULL when z2/mean2/rstd2 are saved", who);
  ME_CHECK(a->xin_mean == nullptr || (a->xin_rstd && a->xin_gamma && a->xin_beta), "%s: incomplete xin_* state", who);
  ME_CHECK((a->xin_mean == nullptr && a->out1_f32 && a->out2_f32) || a->dtype == ME_BF16,
           "%s: the deferred LayerNorm outputs are a ME_BF16 option", who);
  return 0;
}

static void fill_attn(const me_layer_args* a, me_attn_args* t) {
  const int d = a->d, dh = a->d / a->H;
  memset(t, 0, sizeof(*t));
  t->dtype = a->dtype;
  t->impl = a->attn_impl;
  t->flags = a->attn_flags;
  t->B = a->B; t->H = a->H; t->Lq = a->Ls; t->Lk = a->Ls; t->dh = dh; t->max_seq = a->max_seq; t->q_pos0 = 0;
  t->q = a->qkv;
  t->k = offs(static_cast<const void*>(a->qkv), a->dtype, d);
  t->v = offs(static_cast<const void*>(a->qkv), a->dtype, 2 * d);
  t->E = a->E;
  t->q_sb = static_cast<int64_t>(a->Ls) * 3 * d; t->q_sh = dh; t->q_si = 3 * d;
  t->k_sb = t->q_sb; t->k_sh = dh; t->k_sj = 3 * d;
  t->v_sb = t->q_sb; t->v_sh = dh; t->v_sj = 3 * d;
  t->keypad = a->keypad;
  t->keypad_ld = a->Ls;
  t->out = a->attn_o;
  t->o_sb = static_cast<int64_t>(a->Ls) * d; t->o_si = d;
  t->lse = a->lse;
  t->pos_dev = nullptr;
  t->stream = a->stream;
  t->p_tiles = a->attn_p;
  t->m_tiles = a->attn_m;
}

// attention output -> out-projection -> LN1 -> FFN -> LN2 (shared by forward and the decode step)
static int layer_tail(const me_layer_args* a, int M, cudaStream_t st) {
  const int d = a->d, di = a->d_inner, dt = a->dtype;
  const float p = a->training ? a->dropout_p : 0.f;
  const uint64_t s1 = a->seed * 4 + 1, s2 = a->seed * 4 + 2;
  if (linear(dt, a->attn_o, a->Wo, a->proj, M, d, d, d, d, d, 0, 0, false, ME_EPI_BIAS, a->bo, nullptr, nullptr, 0, st))
    return 1;
  if (launch_add_ln_fwd(a->x_f32, a->proj, dt, a->ln1_w, a->ln1_b, a->ln_eps, M, d, p, s1, a->out1_f32, a->out1_T,
                        a->z1, a->mean1, a->rstd1, a->xin_mean, a->xin_rstd, a->xin_gamma, a->xin_beta, st))
    return 1;
  if (linear(dt, a->out1_T, a->W1, a->h, M, di, d, d, d, di, 0, 0, false, ME_EPI_BIAS | ME_EPI_RELU, a->b1, nullptr,
             nullptr, 0, st))
    return 1;
  if (linear(dt, a->h, a->W2, a->proj, M, d, di, di, di, d, 0, 0, false, ME_EPI_BIAS, a->b2, nullptr, nullptr, 0, st))
    return 1;
  // (out1 not materialised in fp32: LN2 re-derives it from LN1's saved state)
  const bool lazy1 = a->out1_f32 == nullptr;
  if (launch_add_ln_fwd(lazy1 ? a->z1 : a->out1_f32, a->proj, dt, a->ln2_w, a->ln2_b, a->ln_eps, M, d, p, s2,
                        a->out2_f32, a->out2_T, a->z2, a->mean2, a->rstd2, lazy1 ? a->mean1 : nullptr,
                        lazy1 ? a->rstd1 : nullptr, lazy1 ? a->ln1_w : nullptr, lazy1 ? a->ln1_b : nullptr, st))
    return 1;
  return 0;
}

// p[k][0 .. n[k]) = 0 for nine small buffers
struct ZeroList {
  float* p[9];
  int64_t n[9];
};
__global__ void zero_list_kernel(ZeroList z) {
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nt = static_cast<int64_t>(gridDim.x) * blockDim.x;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    for (int64_t i = t0; i < z.n[k]; i += nt) z.p[k][i] = 0.f;
}

}  // namespace me

using namespace me;

extern "C" int me_layer_forward(const me_layer_args* a) {
  if (check_layer(a, "me_layer_forward")) return 1;
  ME_CHECK(a->Ls <= a->max_seq, "me_layer_forward: Ls %d > max_seq %d", a->Ls, a->max_seq);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const int d = a->d, M = a->B * a->Ls, dt = a->dtype;
  if (linear(dt, a->x_T, a->Wqkv, a->qkv, M, 3 * d, d, d, d, 3 * d, 0, 0, false, ME_EPI_BIAS, a->bqkv, nullptr,
             nullptr, 0, st))
    return 1;
  me_attn_args t;
  fill_attn(a, &t);
  if (me_attention_forward(&t)) return 1;
  return layer_tail(a, M, st);
}

extern "C" int me_layer_backward(const me_layer_bwd_args* b) {
  ME_CHECK(b != nullptr, "me_layer_backward: NULL args");
  const me_layer_args* a = &b->f;
  if (check_layer(a, "me_layer_backward")) return 1;
  ME_CHECK(a->z1 && a->z2 && a->mean1 && a->rstd1 && a->mean2 && a->rstd2 && a->lse,
           "me_layer_backward: saved z/mean/rstd/lse missing");
  ME_CHECK(b->d_out && b->d_x && b->dWqkv && b->dbqkv && b->dE && b->dWo && b->dbo && b->dln1_w && b->dln1_b &&
               b->dW1 && b->db1 && b->dW2 && b->db2 && b->dln2_w && b->dln2_b,
           "me_layer_backward: NULL gradient pointer");
  ME_CHECK(b->g_a && b->g_b && b->g_T && b->g_h && b->g_qkv && b->g_o && b->dsum, "me_layer_backward: NULL scratch");
  ME_CHECK(a->attn_impl != ME_ATTN_TENSOR || b->attn_ws, "me_layer_backward: attn_ws required for ME_ATTN_TENSOR");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const int d = a->d, di = a->d_inner, M = a->B * a->Ls, dt = a->dtype, H = a->H, dh = d / H;
  const float p = a->training ? a->dropout_p : 0.f;
  const uint64_t s1 = a->seed * 4 + 1, s2 = a->seed * 4 + 2;

  // accumulating outputs start from zero (one launch: nine memsets per layer were ~100 launches per step)
  {
    ZeroList z;
    float* ptrs[9] = {b->dln2_w, b->dln2_b, b->dln1_w, b->dln1_b, b->db2, b->db1, b->dbo, b->dbqkv, b->dE};
    const int64_t counts[9] = {d, d, d, d, d, di, d, 3 * d, static_cast<int64_t>(a->max_seq) * dh};
    for (int k = 0; k < 9; ++k) {
      z.p[k] = ptrs[k];
      z.n[k] = counts[k];
    }
    zero_list_kernel<<<64, 256, 0, st>>>(z);
    ME_LAUNCH_CHECK();
  }

  // ---- FFN block: out2 = LN2(out1 + drop(W2 relu(W1 out1 + b1) + b2))
  // bf16 path: the input gradients of the two sub-layers leave their GEMMs in bf16 (the fast TMA-store
  // epilogue; the same rounding point as the reference under autocast) and the following LayerNorm backward
  // adds them to the fp32 residual-stream gradient itself -- no fp32 [M, d] round trip through the GEMM epilogue.
  const bool split = dt == ME_BF16 && b->d_x_T != nullptr;
  if (launch_add_ln_bwd(b->d_out, nullptr, dt == ME_BF16 ? b->d_out_T : nullptr, a->z2, a->mean2, a->rstd2, a->ln2_w, M, d,
                        p, s2, dt, b->g_a, b->g_T, b->dln2_w, b->dln2_b, b->db2, st))  // db2 = column sums of the masked gradient
    return 1;
  // dW2[d, di] = g_T^T . h
  if (linear(dt, b->g_T, a->h, b->dW2, d, di, M, d, di, di, 1, 1, true, 0, nullptr, nullptr, nullptr, 0, st)) return 1;
  // g_h[M, di] = (g_T . W2) masked by relu; db1 = its column sums, taken from the staged output tiles by the GEMM's
  // epilogue when that path runs (bf16), else by a pass over g_h
  bool db1_done = false;
  if (linear(dt, b->g_T, a->W2, b->g_h, M, di, d, d, di, di, 0, 1, false, ME_EPI_RELU_MASK, nullptr, nullptr, a->h,
             di, st, b->db1, &db1_done))
    return 1;
  // (the attention workspace is idle outside me_attention_backward: scratch for the partial column sums)
  const int64_t ws_floats = b->attn_ws ? me_attention_backward_workspace_floats(a->B, H, a->Ls, dh, a->max_seq) : 0;
  if (!db1_done && launch_colsum_ws(b->g_h, dt, M, di, di, b->db1, b->attn_ws, ws_floats, st)) return 1;
  // dW1[di, d] = g_h^T . out1
  if (linear(dt, b->g_h, a->out1_T, b->dW1, di, d, M, di, d, d, 1, 1, true, 0, nullptr, nullptr, nullptr, 0, st))
    return 1;
  if (split) {
    // g_o (free until the out-projection dgrad) <- g_h . W1 in bf16; LN1 backward reads g_a + g_o and writes the
    // fp32 part of d_x directly
    if (linear(dt, b->g_h, a->W1, b->g_o, M, d, di, di, d, d, 0, 1, false, 0, nullptr, nullptr, nullptr, 0, st)) return 1;
    if (launch_add_ln_bwd(b->g_a, nullptr, b->g_o, a->z1, a->mean1, a->rstd1, a->ln1_w, M, d, p, s1, dt, b->d_x, b->g_T,
                          b->dln1_w, b->dln1_b, b->dbo, st))
      return 1;
  } else {
    // g_b[M, d] = g_h . W1 + g_a   (total gradient w.r.t. out1, fp32)
    if (linear(dt, b->g_h, a->W1, b->g_b, M, d, di, di, d, d, 0, 1, true, ME_EPI_ADD_F32, nullptr, b->g_a, nullptr, 0,
               st))
      return 1;
    // ---- attention block: out1 = LN1(x + drop(Wo attn + bo))
    if (launch_add_ln_bwd(b->g_b, nullptr, nullptr, a->z1, a->mean1, a->rstd1, a->ln1_w, M, d, p, s1, dt, b->g_a, b->g_T,
                          b->dln1_w, b->dln1_b, b->dbo, st))
      return 1;
  }
  if (linear(dt, b->g_T, a->attn_o, b->dWo, d, d, M, d, d, d, 1, 1, true, 0, nullptr, nullptr, nullptr, 0, st))
    return 1;
  if (linear(dt, b->g_T, a->Wo, b->g_o, M, d, d, d, d, d, 0, 1, false, 0, nullptr, nullptr, nullptr, 0, st)) return 1;

  me_attn_bwd_args t;
  memset(&t, 0, sizeof(t));
  fill_attn(a, &t.f);
  t.dout = b->g_o;
  t.dq = b->g_qkv;
  t.dk = offs(b->g_qkv, dt, d);
  t.dv = offs(b->g_qkv, dt, 2 * d);
  t.dE = b->dE;
  t.dsum = b->dsum;
  t.dq_acc = b->attn_ws;
  if (me_attention_backward(&t)) return 1;

  if (launch_colsum_ws(b->g_qkv, dt, M, 3 * d, 3 * d, b->dbqkv, b->attn_ws, ws_floats, st)) return 1;
  // dWqkv[3d, d] = g_qkv^T . x
  if (linear(dt, b->g_qkv, a->x_T, b->dWqkv, 3 * d, d, M, 3 * d, d, d, 1, 1, true, 0, nullptr, nullptr, nullptr, 0,
             st))
    return 1;
  if (split) {
    // d_x = (fp32 part, written by LN1 backward above) + d_x_T, with d_x_T = g_qkv . Wqkv in bf16
    if (linear(dt, b->g_qkv, a->Wqkv, b->d_x_T, M, d, 3 * d, 3 * d, d, d, 0, 1, false, 0, nullptr, nullptr, nullptr, 0, st))
      return 1;
  } else {
    // d_x[M, d] = g_qkv . Wqkv + g_a
    if (linear(dt, b->g_qkv, a->Wqkv, b->d_x, M, d, 3 * d, 3 * d, d, d, 0, 1, true, ME_EPI_ADD_F32, nullptr, b->g_a,
               nullptr, 0, st))
      return 1;
  }
  return 0;
}

extern "C" int me_decode_layer_step(const me_decode_layer_args* c) {
  ME_CHECK(c != nullptr, "me_decode_layer_step: NULL args");
  const me_layer_args* a = &c->f;
  if (check_layer(a, "me_decode_layer_step")) return 1;
  ME_CHECK(a->Ls == 1, "me_decode_layer_step: Ls must be 1");
  ME_CHECK(c->k_cache && c->v_cache && c->t_dev && c->T_max > 0 && c->T_max <= a->max_seq,
           "me_decode_layer_step: bad cache arguments");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const int d = a->d, B = a->B, H = a->H, dh = d / H, dt = a->dtype;
  if (linear(dt, a->x_T, a->Wqkv, a->qkv, B, 3 * d, d, d, d, 3 * d, 0, 0, false, ME_EPI_BIAS, a->bqkv, nullptr,
             nullptr, 0, st))
    return 1;
  if (launch_kv_write(a->qkv, dt, B, 1, H, dh, c->k_cache, c->v_cache, c->T_max, 0, c->t_dev, st)) return 1;
  me_attn_args t;
  memset(&t, 0, sizeof(t));
  t.dtype = dt;
  t.impl = ME_ATTN_SIMT;  // one query row per sequence: HBM-bound streaming of the cache
  t.B = B; t.H = H; t.Lq = 1; t.Lk = 0; t.dh = dh; t.max_seq = a->max_seq; t.q_pos0 = 0;
  t.q = a->qkv; t.k = c->k_cache; t.v = c->v_cache; t.E = a->E;
  t.q_sb = 3 * d; t.q_sh = dh; t.q_si = 0;
  t.k_sb = static_cast<int64_t>(H) * c->T_max * dh; t.k_sh = static_cast<int64_t>(c->T_max) * dh; t.k_sj = dh;
  t.v_sb = t.k_sb; t.v_sh = t.k_sh; t.v_sj = dh;
  t.keypad = a->keypad; t.keypad_ld = c->T_max;
  t.out = a->attn_o; t.o_sb = d; t.o_si = 0;
  t.lse = nullptr;
  t.pos_dev = c->t_dev;
  t.stream = a->stream;
  if (me_attention_forward(&t)) return 1;
  return layer_tail(a, B, st);
}
