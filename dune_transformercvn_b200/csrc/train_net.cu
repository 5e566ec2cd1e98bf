// Token assembly + post-norm transformer encoder + event / prong heads, TRAIN mode: forward with batch-statistics
// BatchNorm1d and dropout, and the hand-written backward (fp32, primitives of train.cu / train_seq.cu).
//
// Reference under autograd in .train():
//   tokens   transformercvn/network/networks/neutrino_full_base_network.py:99-125 (LinearBlock = Linear(no bias) ->
//            BatchNorm1d over the B+T packed rows -> PReLU -> Dropout, layers/prong_feature_embedding.py:7-33)
//   encoder  layers/prong_custom_bert_encoder.py:45-75: torch TransformerEncoderLayer, post-norm, gelu(erf); dropout
//            on the attention probabilities, after out_proj, after gelu and after linear2
//   heads    layers/prong_decoder.py:13-16; layers/prong_target_decoder.py:19-41 (BatchNorm1d over ALL L*B rows,
//            padded slots included)
// Token rows: X[R = B*S][D], row = event*S + slot, slot 0 = event token.  Prong-head rows are in (slot, event)
// order like the reference's (L,B,D) -> (L*B,D) reshape; prong logits are returned in that order, the caller
// exposes them as (B,L,C) by a transposed view exactly like neutrino_full_base_network.py:188.
// Linear weights stay in the reference's [out][in] layout: that IS the [K][N] operand of the input-gradient
// GEMM, and the weight-gradient kernel writes [out][in] directly when fed (dY, X); only the forward GEMM needs
// the transposed copy, rebuilt per step.
#include "kernels.h"

namespace tcvn {

namespace {

constexpr int NOGAP = 1 << 30;
constexpr int kMaxLayers = 16;

struct EncArena { int64_t wqkv, bqkv, wo, bo, w1, b1, w2, b2, ln1w, ln1b, ln2w, ln2b; };
struct DecArena { int64_t w, b, bn_w, bn_b, bn_rm, bn_rv, alpha; int cin, cout; };

struct NetPlan {
  tcvn_seq_desc d;
  int B, L, T, S, R, D, in_dim, rows_tok, LB;
  // arenas
  int64_t c_w, c_bn_w, c_bn_b, c_rm, c_rv, c_alpha;
  EncArena enc[kMaxLayers];
  DecArena dec[TCVN_MAX_DECODER_LAYERS];
  int64_t o_w, o_b, dec_last_cin;
  // workspace (bytes)
  size_t offsets, smask, xtok, rows_pre, fold_c, rows, x0, hidden, hrows, sums, cw_t, ew_t, ow_t;
  struct LayerWs { size_t qkv, P, ctx, pre1, st1, x1, h, hg, pre2, st2, x2, wqkv_t, wo_t, w1_t, w2_t; } lw[kMaxLayers];
  struct DecWs { size_t z, fold, a, w_t; } dw[TCVN_MAX_DECODER_LAYERS];
  // backward scratch
  size_t g0, g1, g2, g3, gq, dhid, dparts;
  size_t bytes;

  static bool build(const tcvn_seq_desc& d, int B, int L, int T, NetPlan* P) {
    if (d.layers < 0 || d.layers > kMaxLayers || d.hidden < 1 || d.hidden > 128 || d.hidden % d.heads) return false;
    if (d.num_decoder_layers < 0 || d.num_decoder_layers > TCVN_MAX_DECODER_LAYERS) return false;
    if (B < 1 || L < 1 || T < 0 || 1 + L > 32 || d.hidden / d.heads > 16) return false;
    P->d = d; P->B = B; P->L = L; P->T = T; P->S = 1 + L; P->R = B * (1 + L); P->D = d.hidden;
    P->in_dim = d.feature_dim + d.pixel_dim + d.position_dim;
    P->rows_tok = B + T; P->LB = L * B;
    const int D = d.hidden, F = d.ffn;
    // ---- arenas (reference state_dict order)
    P->c_w = 0; P->c_bn_w = (int64_t)D * P->in_dim; P->c_bn_b = P->c_bn_w + D; P->c_rm = P->c_bn_b + D;
    P->c_rv = P->c_rm + D; P->c_alpha = P->c_rv + D;
    int64_t a = 0;
    for (int l = 0; l < d.layers; ++l) {
      EncArena& E = P->enc[l];
      E.wqkv = a; a += (int64_t)3 * D * D; E.bqkv = a; a += 3 * D;
      E.wo = a; a += (int64_t)D * D; E.bo = a; a += D;
      E.w1 = a; a += (int64_t)F * D; E.b1 = a; a += F;
      E.w2 = a; a += (int64_t)D * F; E.b2 = a; a += D;
      E.ln1w = a; a += D; E.ln1b = a; a += D; E.ln2w = a; a += D; E.ln2b = a; a += D;
    }
    a = 0;
    int cin = D;
    for (int i = 0; i < d.num_decoder_layers; ++i) {
      DecArena& X = P->dec[i];
      X.cin = cin; X.cout = d.decoder_widths[i];
      if (X.cout < 4 || X.cout % 4 || X.cout > D) return false;
      X.w = a; a += (int64_t)X.cout * cin; X.b = a; a += X.cout;
      X.bn_w = a; a += X.cout; X.bn_b = a; a += X.cout; X.bn_rm = a; a += X.cout; X.bn_rv = a; a += X.cout;
      X.alpha = a; a += X.cout;
      cin = X.cout;
    }
    P->dec_last_cin = cin;
    P->o_w = a; P->o_b = a + (int64_t)d.num_prong_classes * cin;
    if (d.num_event_classes % 4 || d.num_prong_classes % 4) return false;
    // ---- workspace
    size_t w = 0;
    auto take = [&](size_t floats) { size_t o = w; w += (floats * 4 + 255) / 256 * 256; return o; };
    const size_t R = P->R, RT = P->rows_tok, LB = P->LB;
    P->offsets = take(B + 1);
    P->smask = take((R + 3) / 4);
    P->xtok = take(RT * P->in_dim);
    P->rows_pre = take(RT * D);
    P->fold_c = take(5 * D);
    P->rows = take(RT * D);
    P->x0 = take(R * D);
    P->hidden = take(R * D);
    P->hrows = take(LB * D);
    P->sums = take(2 * 3 * (size_t)(3 * D > P->in_dim ? 3 * D : P->in_dim));
    P->cw_t = take((size_t)P->in_dim * D);
    P->ew_t = take((size_t)D * d.num_event_classes);
    P->ow_t = take((size_t)cin * d.num_prong_classes);
    for (int l = 0; l < d.layers; ++l) {
      auto& X = P->lw[l];
      X.qkv = take(R * 3 * D); X.P = take((size_t)B * d.heads * P->S * P->S); X.ctx = take(R * D);
      X.pre1 = take(R * D); X.st1 = take(2 * R); X.x1 = take(R * D);
      X.h = take(R * F); X.hg = take(R * F);
      X.pre2 = take(R * D); X.st2 = take(2 * R); X.x2 = take(R * D);
      X.wqkv_t = take((size_t)3 * D * D); X.wo_t = take((size_t)D * D); X.w1_t = take((size_t)D * F); X.w2_t = take((size_t)F * D);
    }
    for (int i = 0; i < d.num_decoder_layers; ++i) {
      auto& X = P->dw[i];
      const DecArena& A = P->dec[i];
      X.z = take(LB * A.cout); X.fold = take(5 * A.cout); X.a = take(LB * A.cout); X.w_t = take((size_t)A.cin * A.cout);
    }
    const size_t widest = (size_t)(3 * D > P->in_dim ? 3 * D : P->in_dim);
    const size_t big = (R > RT ? R : RT) > LB ? (R > RT ? R : RT) : LB;
    P->g0 = take(big * widest); P->g1 = take(big * widest); P->g2 = take(big * widest); P->g3 = take(big * widest);
    P->gq = take(R * 3 * D);
    P->dhid = take(R * D);
    P->dparts = take(colsum_parts_bytes() / 4);   // per-slab partial sums of the column reductions
    P->bytes = w;
    return true;
  }
};

__global__ void seq_mask_kernel(const uint8_t* __restrict__ event_mask, const uint8_t* __restrict__ prong_mask, int B, int L,
                                uint8_t* __restrict__ smask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int S = 1 + L;
  if (i >= B * S) return;
  const int b = i / S, s = i - b * S;
  smask[i] = s == 0 ? (event_mask ? (event_mask[b] != 0) : 1) : (prong_mask[(size_t)b * L + s - 1] != 0);
}

// BatchNorm1d batch statistics -> fold; running update (see train_cnn.cu for the mapped variant)
struct Fin1 {
  const double* sums; int C; double count;
  const float *gamma, *beta, *alpha;
  float eps, momentum;
  float *rm, *rv, *fold;
};
__global__ void bn1d_finalize_kernel(const Fin1 a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const double mean = a.sums[c] / a.count;
  double var = a.sums[a.C + c] / a.count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rs = (float)(1.0 / sqrt(var + (double)a.eps));
  const float sc = a.gamma[c] * rs;
  a.fold[c] = sc;
  a.fold[a.C + c] = a.beta[c] - (float)mean * sc;
  a.fold[2 * a.C + c] = a.alpha[c];
  a.fold[3 * a.C + c] = (float)mean;
  a.fold[4 * a.C + c] = rs;
  if (a.rm) {
    const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
    a.rm[c] = (1.f - a.momentum) * a.rm[c] + a.momentum * (float)mean;
    a.rv[c] = (1.f - a.momentum) * a.rv[c] + a.momentum * (float)unbiased;
  }
}

__global__ void add_sums1_kernel(const double* __restrict__ sums, int c, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) dst[i] += (float)sums[i];
}

struct NWalk {
  const NetPlan& P;
  float *position, *combined, *encoder, *event_dec, *prong_dec;              // parameter arenas
  float *g_position, *g_combined, *g_encoder, *g_event_dec, *g_prong_dec;    // gradient arenas (backward only)
  const uint8_t* prong_mask;
  char* ws;
  cudaStream_t st;
  float p_drop, momentum;
  uint64_t seed;

  float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  double* sums() const { return reinterpret_cast<double*>(ws + P.sums); }
  double* dparts() const { return reinterpret_cast<double*>(ws + P.dparts); }
  uint8_t* smask() const { return reinterpret_cast<uint8_t*>(ws + P.smask); }
  int* offsets() const { return reinterpret_cast<int*>(ws + P.offsets); }
  uint64_t sid(int k) const { return 3ull * 4096 + (uint64_t)k; }

  int transpose(const float* src, int n_out, int k_in, float* dst) {
    return repack(src, n_out, k_in, 1, NOGAP, NOGAP, k_in, n_out, false, false, dst, st);
  }
  // Y[rows, N] = X[rows, K] W[K][N] + bias
  int lin(const float* X, int ldx, long long rows, int K, const float* W, int N, const float* bias, float* Y, int ldy) {
    return tcvn_t_gemm(X, ldx, rows, K, 1, nullptr, W, N, nullptr, 0, 0, bias, Y, ldy, 0, 0, 0, 0, st);
  }
  int bias_grad(const float* G, int ldg, int col0, int C, long long rows, float* dst) {
    TCVN_TRY(colsums_into(2, G, ldg, col0, nullptr, 0, 0, nullptr, C, rows, 0, 0, sums(), C, dparts(), st));
    add_sums1_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums(), C, dst);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }
  // gradients of Y = X W^T + b with W [n_out][k_in] in the reference layout: dW += dY^T X, db += colsum dY,
  // dX = dY W  (dX nullable)
  int lin_bwd(const float* dY, int lddy, long long rows, int n_out, const float* X, int ldx, int k_in, const float* W,
              float* dW, float* db, float* dX, int lddx) {
    if (db) TCVN_TRY(bias_grad(dY, lddy, 0, n_out, rows, db));
    TCVN_TRY(wgrad_f32(dY, lddy, rows, n_out, X, ldx, k_in, dW, reinterpret_cast<float*>(dparts()), colsum_parts_bytes() / 4, st));
    if (dX) TCVN_TRY(tcvn_t_gemm(dY, lddy, rows, n_out, 1, nullptr, W, k_in, nullptr, 0, 0, nullptr, dX, lddx, 0, 0, 0, 0, st));
    return TCVN_OK;
  }
  int bn_fwd(const float* Z, int C, long long rows, const float* gamma, const float* beta, const float* alpha, float* rm,
             float* rv, float* fold, float* out) {
    TCVN_TRY(colsums_into(0, Z, C, 0, nullptr, 0, 0, nullptr, C, rows, 0, 0, sums(), C, dparts(), st));
    Fin1 a;
    a.sums = sums(); a.C = C; a.count = (double)rows; a.gamma = gamma; a.beta = beta; a.alpha = alpha; a.eps = P.d.bn_eps;
    a.momentum = momentum; a.rm = rm; a.rv = rv; a.fold = fold;
    bn1d_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(a);
    TCVN_LAUNCH_CHECK();
    return tcvn_t_bnact_fwd(Z, C, 0, fold, C, rows, 0, 0, out, C, 0, st);
  }
  // in place: D <- gradient w.r.t. the BatchNorm input; parameter gradients accumulated
  int bn_bwd(const float* Z, float* D, int C, long long rows, const float* fold, float* dgamma, float* dbeta, float* dalpha) {
    TCVN_TRY(colsums_into(1, Z, C, 0, D, C, 0, fold, C, rows, 0, 0, sums(), C, dparts(), st));
    return tcvn_t_bnact_bwd_apply(D, C, 0, Z, C, 0, fold, sums(), C, (double)rows, D, C, 0, 0, rows, 0, 0, dgamma, dbeta, dalpha,
                                  st);
  }
  int dropout(float* X, int C, long long rows, int k) { return tcvn_t_dropout(X, C, 0, C, rows, seed, sid(k), p_drop, st); }
  int copy(float* dst, const float* src, size_t floats) {
    TCVN_CUDA(cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return TCVN_OK;
  }

  int forward(const float* ev_emb, const float* pr_emb, const uint8_t* event_mask, float* ev_logits, float* pr_logits) {
    const tcvn_seq_desc& d = P.d;
    const int B = P.B, L = P.L, T = P.T, S = P.S, R = P.R, D = P.D, F = d.ffn, RT = P.rows_tok, LB = P.LB;
    // ---- forward weight layouts
    TCVN_TRY(transpose(combined + P.c_w, D, P.in_dim, f(P.cw_t)));
    for (int l = 0; l < d.layers; ++l) {
      const EncArena& E = P.enc[l];
      TCVN_TRY(transpose(encoder + E.wqkv, 3 * D, D, f(P.lw[l].wqkv_t)));
      TCVN_TRY(transpose(encoder + E.wo, D, D, f(P.lw[l].wo_t)));
      TCVN_TRY(transpose(encoder + E.w1, F, D, f(P.lw[l].w1_t)));
      TCVN_TRY(transpose(encoder + E.w2, D, F, f(P.lw[l].w2_t)));
    }
    TCVN_TRY(transpose(event_dec, d.num_event_classes, D, f(P.ew_t)));
    for (int i = 0; i < d.num_decoder_layers; ++i)
      TCVN_TRY(transpose(prong_dec + P.dec[i].w, P.dec[i].cout, P.dec[i].cin, f(P.dw[i].w_t)));
    TCVN_TRY(transpose(prong_dec + P.o_w, d.num_prong_classes, (int)P.dec_last_cin, f(P.ow_t)));
    // ---- tokens
    TCVN_TRY(tcvn_t_tokens(0, 0, nullptr, nullptr, nullptr, prong_mask, offsets(), B, T, L, D, 0, 0, 0, nullptr, st));
    seq_mask_kernel<<<ceil_div(R, 256), 256, 0, st>>>(event_mask, prong_mask, B, L, smask());
    TCVN_LAUNCH_CHECK();
    TCVN_TRY(tcvn_t_tokens(1, 0, const_cast<float*>(ev_emb), const_cast<float*>(pr_emb), position, prong_mask, offsets(), B, T, L,
                           D, d.pixel_dim, d.feature_dim, d.position_dim, f(P.xtok), st));
    TCVN_TRY(lin(f(P.xtok), P.in_dim, RT, P.in_dim, f(P.cw_t), D, nullptr, f(P.rows_pre), D));
    TCVN_TRY(bn_fwd(f(P.rows_pre), D, RT, combined + P.c_bn_w, combined + P.c_bn_b, combined + P.c_alpha, combined + P.c_rm,
                    combined + P.c_rv, f(P.fold_c), f(P.rows)));
    TCVN_TRY(dropout(f(P.rows), D, RT, 0));
    TCVN_TRY(tcvn_t_tokens(2, 0, f(P.rows), nullptr, nullptr, prong_mask, offsets(), B, T, L, D, 0, 0, 0, f(P.x0), st));
    TCVN_TRY(tcvn_t_eltwise(3, f(P.x0), nullptr, smask(), D, (long long)R * D, f(P.x0), st));
    // ---- encoder
    const float* x = f(P.x0);
    for (int l = 0; l < d.layers; ++l) {
      const EncArena& E = P.enc[l];
      const auto& W = P.lw[l];
      float* t0 = f(P.g0);
      TCVN_TRY(lin(x, D, R, D, f(W.wqkv_t), 3 * D, encoder + E.bqkv, f(W.qkv), 3 * D));
      TCVN_TRY(tcvn_t_attention(0, f(W.qkv), smask(), B, S, d.heads, D, f(W.P), f(W.ctx), nullptr, p_drop, seed, sid(16 + 8 * l),
                                st));
      TCVN_TRY(lin(f(W.ctx), D, R, D, f(W.wo_t), D, encoder + E.bo, t0, D));
      TCVN_TRY(dropout(t0, D, R, 16 + 8 * l + 1));
      TCVN_TRY(tcvn_t_layernorm(0, x, t0, D, R, encoder + E.ln1w, encoder + E.ln1b, d.ln_eps, f(W.pre1), f(W.st1), f(W.x1), nullptr,
                                nullptr, st));
      TCVN_TRY(lin(f(W.x1), D, R, D, f(W.w1_t), F, encoder + E.b1, f(W.h), F));
      TCVN_TRY(tcvn_t_eltwise(0, f(W.h), nullptr, nullptr, F, (long long)R * F, f(W.hg), st));
      TCVN_TRY(dropout(f(W.hg), F, R, 16 + 8 * l + 2));
      TCVN_TRY(lin(f(W.hg), F, R, F, f(W.w2_t), D, encoder + E.b2, t0, D));
      TCVN_TRY(dropout(t0, D, R, 16 + 8 * l + 3));
      TCVN_TRY(tcvn_t_layernorm(0, f(W.x1), t0, D, R, encoder + E.ln2w, encoder + E.ln2b, d.ln_eps, f(W.pre2), f(W.st2), f(W.x2),
                                nullptr, nullptr, st));
      x = f(W.x2);
    }
    TCVN_TRY(tcvn_t_eltwise(3, x, nullptr, smask(), D, (long long)R * D, f(P.hidden), st));
    // ---- heads
    TCVN_TRY(lin(f(P.hidden), S * D, B, D, f(P.ew_t), d.num_event_classes, event_dec + (int64_t)d.num_event_classes * D,
                 ev_logits, d.num_event_classes));
    TCVN_TRY(tcvn_t_tokens(3, 0, f(P.hidden), f(P.hrows), nullptr, prong_mask, offsets(), B, T, L, D, 0, 0, 0, nullptr, st));
    const float* a = f(P.hrows);
    for (int i = 0; i < d.num_decoder_layers; ++i) {
      const DecArena& A = P.dec[i];
      const auto& W = P.dw[i];
      TCVN_TRY(lin(a, A.cin, LB, A.cin, f(W.w_t), A.cout, prong_dec + A.b, f(W.z), A.cout));
      TCVN_TRY(bn_fwd(f(W.z), A.cout, LB, prong_dec + A.bn_w, prong_dec + A.bn_b, prong_dec + A.alpha, prong_dec + A.bn_rm,
                      prong_dec + A.bn_rv, f(W.fold), f(W.a)));
      TCVN_TRY(dropout(f(W.a), A.cout, LB, 1 + i));
      a = f(W.a);
    }
    TCVN_TRY(lin(a, (int)P.dec_last_cin, LB, (int)P.dec_last_cin, f(P.ow_t), d.num_prong_classes, prong_dec + P.o_b, pr_logits,
                 d.num_prong_classes));
    return TCVN_OK;
  }

  int backward(float* d_ev_logits, float* d_pr_logits, float* d_ev_emb, float* d_pr_emb) {
    const tcvn_seq_desc& d = P.d;
    const int B = P.B, L = P.L, T = P.T, S = P.S, R = P.R, D = P.D, F = d.ffn, RT = P.rows_tok, LB = P.LB;
    const int E = d.num_event_classes, C = d.num_prong_classes;
    float *g0 = f(P.g0), *g1 = f(P.g1), *g2 = f(P.g2), *g3 = f(P.g3);
    // ---- prong head
    const int nd = d.num_decoder_layers;
    const float* a_last = nd ? f(P.dw[nd - 1].a) : f(P.hrows);
    TCVN_TRY(lin_bwd(d_pr_logits, C, LB, C, a_last, (int)P.dec_last_cin, (int)P.dec_last_cin, prong_dec + P.o_w,
                     g_prong_dec + P.o_w, g_prong_dec + P.o_b, g0, (int)P.dec_last_cin));
    float* cur = g0;
    float* other = g1;
    for (int i = nd - 1; i >= 0; --i) {
      const DecArena& A = P.dec[i];
      const auto& W = P.dw[i];
      TCVN_TRY(dropout(cur, A.cout, LB, 1 + i));
      TCVN_TRY(bn_bwd(f(W.z), cur, A.cout, LB, f(W.fold), g_prong_dec + A.bn_w, g_prong_dec + A.bn_b, g_prong_dec + A.alpha));
      const float* prev = i ? f(P.dw[i - 1].a) : f(P.hrows);
      TCVN_TRY(lin_bwd(cur, A.cout, LB, A.cout, prev, A.cin, A.cin, prong_dec + A.w, g_prong_dec + A.w, g_prong_dec + A.b, other,
                       A.cin));
      float* t = cur; cur = other; other = t;
    }
    // cur = gradient of the prong-head rows [LB, D]
    float* dhid = f(P.dhid);
    TCVN_CUDA(cudaMemsetAsync(dhid, 0, sizeof(float) * (size_t)R * D, st));
    TCVN_TRY(lin_bwd(d_ev_logits, E, B, E, f(P.hidden), S * D, D, event_dec, g_event_dec, g_event_dec + (int64_t)E * D, dhid,
                     S * D));
    TCVN_TRY(tcvn_t_tokens(3, 1, dhid, cur, nullptr, prong_mask, offsets(), B, T, L, D, 0, 0, 0, nullptr, st));
    TCVN_TRY(tcvn_t_eltwise(3, dhid, nullptr, smask(), D, (long long)R * D, dhid, st));
    // ---- encoder, last layer first.  dx = gradient of the layer output
    float* dx = dhid;
    for (int l = d.layers - 1; l >= 0; --l) {
      const EncArena& A = P.enc[l];
      const auto& W = P.lw[l];
      const float* x_in = l ? f(P.lw[l - 1].x2) : f(P.x0);
      // LN2
      TCVN_TRY(tcvn_t_layernorm(1, dx, nullptr, D, R, encoder + A.ln2w, encoder + A.ln2b, d.ln_eps, f(W.pre2), f(W.st2), g0,
                                g_encoder + A.ln2w, g_encoder + A.ln2b, st));      // g0 = d(x1 + ff)
      TCVN_TRY(copy(g1, g0, (size_t)R * D));
      TCVN_TRY(dropout(g1, D, R, 16 + 8 * l + 3));                                  // g1 = d ff
      TCVN_TRY(lin_bwd(g1, D, R, D, f(W.hg), F, F, encoder + A.w2, g_encoder + A.w2, g_encoder + A.b2, g2, F));  // g2 = d hg
      TCVN_TRY(dropout(g2, F, R, 16 + 8 * l + 2));
      TCVN_TRY(tcvn_t_eltwise(1, f(W.h), g2, nullptr, F, (long long)R * F, g2, st));                              // g2 = d h
      TCVN_TRY(lin_bwd(g2, F, R, F, f(W.x1), D, D, encoder + A.w1, g_encoder + A.w1, g_encoder + A.b1, g1, D));  // g1 = d x1 (ffn)
      TCVN_TRY(tcvn_t_eltwise(2, g1, g0, nullptr, D, (long long)R * D, g1, st));                                  // + residual
      // LN1
      TCVN_TRY(tcvn_t_layernorm(1, g1, nullptr, D, R, encoder + A.ln1w, encoder + A.ln1b, d.ln_eps, f(W.pre1), f(W.st1), g0,
                                g_encoder + A.ln1w, g_encoder + A.ln1b, st));      // g0 = d(x + attn_out)
      TCVN_TRY(copy(g1, g0, (size_t)R * D));
      TCVN_TRY(dropout(g1, D, R, 16 + 8 * l + 1));                                  // g1 = d attn_out
      TCVN_TRY(lin_bwd(g1, D, R, D, f(W.ctx), D, D, encoder + A.wo, g_encoder + A.wo, g_encoder + A.bo, g2, D));  // g2 = d ctx
      TCVN_TRY(tcvn_t_attention(1, f(W.qkv), smask(), B, S, d.heads, D, f(W.P), g2, f(P.gq), p_drop, seed, sid(16 + 8 * l), st));
      TCVN_TRY(lin_bwd(f(P.gq), 3 * D, R, 3 * D, x_in, D, D, encoder + A.wqkv, g_encoder + A.wqkv, g_encoder + A.bqkv, g3, D));
      TCVN_TRY(tcvn_t_eltwise(2, g3, g0, nullptr, D, (long long)R * D, g3, st));   // g3 = d x_in
      TCVN_TRY(copy(dhid, g3, (size_t)R * D));
      dx = dhid;
    }
    TCVN_TRY(tcvn_t_eltwise(3, dx, nullptr, smask(), D, (long long)R * D, dx, st));
    // ---- tokens
    TCVN_TRY(tcvn_t_tokens(2, 1, g0, nullptr, nullptr, prong_mask, offsets(), B, T, L, D, 0, 0, 0, dx, st));   // g0 = d rows [RT, D]
    TCVN_TRY(dropout(g0, D, RT, 0));
    TCVN_TRY(bn_bwd(f(P.rows_pre), g0, D, RT, f(P.fold_c), g_combined + P.c_bn_w, g_combined + P.c_bn_b, g_combined + P.c_alpha));
    TCVN_TRY(lin_bwd(g0, D, RT, D, f(P.xtok), P.in_dim, P.in_dim, combined + P.c_w, g_combined + P.c_w, nullptr, g1, P.in_dim));
    TCVN_TRY(bias_grad(g1, P.in_dim, d.pixel_dim + d.feature_dim, d.position_dim, RT, g_position));
    TCVN_TRY(tcvn_t_tokens(1, 1, d_ev_emb, d_pr_emb, position, prong_mask, offsets(), B, T, L, D, d.pixel_dim, d.feature_dim,
                           d.position_dim, g1, st));
    return TCVN_OK;
  }
};

}  // namespace
}  // namespace tcvn

using namespace tcvn;

extern "C" size_t tcvn_seq_train_workspace_bytes(const tcvn_seq_desc* d, int n_events, int max_prongs, int n_prongs) {
  NetPlan P;
  if (!d || !NetPlan::build(*d, n_events, max_prongs, n_prongs, &P)) { set_error("seq_train: bad descriptor or sizes"); return 0; }
  return P.bytes;
}

extern "C" int tcvn_seq_train_forward(const tcvn_seq_desc* d, const float* position, float* combined, const float* encoder,
                                      const float* event_decoder, float* prong_decoder, const float* event_embedding,
                                      const float* prong_embedding, const uint8_t* event_mask, const uint8_t* prong_mask,
                                      int n_events, int max_prongs, int n_prongs, float p_drop, float momentum, uint64_t seed,
                                      float* event_logits, float* prong_logits, void* workspace, size_t workspace_bytes,
                                      tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && position && combined && encoder && event_decoder && prong_decoder && event_embedding && prong_mask &&
                     event_logits && prong_logits && workspace && (prong_embedding || n_prongs == 0),
                 "seq_train_forward: null pointer");
  NetPlan P;
  TCVN_CHECK_ARG(NetPlan::build(*d, n_events, max_prongs, n_prongs, &P), "seq_train_forward: bad descriptor or sizes");
  TCVN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "seq_train_forward: dropout probability out of range");
  if (workspace_bytes < P.bytes)
    return fail(TCVN_ERR_WORKSPACE, "seq_train_forward: workspace %zu < %zu bytes", workspace_bytes, P.bytes);
  NWalk w{P, const_cast<float*>(position), combined, const_cast<float*>(encoder), const_cast<float*>(event_decoder),
          prong_decoder, nullptr, nullptr, nullptr, nullptr, nullptr, prong_mask, static_cast<char*>(workspace), stream, p_drop,
          momentum, seed};
  return w.forward(event_embedding, prong_embedding, event_mask, event_logits, prong_logits);
}

extern "C" int tcvn_seq_train_backward(const tcvn_seq_desc* d, const float* position, const float* combined,
                                       const float* encoder, const float* event_decoder, const float* prong_decoder,
                                       float* g_position, float* g_combined, float* g_encoder, float* g_event_decoder,
                                       float* g_prong_decoder, const uint8_t* prong_mask, int n_events, int max_prongs,
                                       int n_prongs, float p_drop, uint64_t seed, float* d_event_logits, float* d_prong_logits,
                                       float* d_event_embedding, float* d_prong_embedding, void* workspace,
                                       size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && position && combined && encoder && event_decoder && prong_decoder && g_position && g_combined &&
                     g_encoder && g_event_decoder && g_prong_decoder && prong_mask && d_event_logits && d_prong_logits &&
                     d_event_embedding && workspace && (d_prong_embedding || n_prongs == 0),
                 "seq_train_backward: null pointer");
  NetPlan P;
  TCVN_CHECK_ARG(NetPlan::build(*d, n_events, max_prongs, n_prongs, &P), "seq_train_backward: bad descriptor or sizes");
  if (workspace_bytes < P.bytes)
    return fail(TCVN_ERR_WORKSPACE, "seq_train_backward: workspace %zu < %zu bytes", workspace_bytes, P.bytes);
  NWalk w{P, const_cast<float*>(position), const_cast<float*>(combined), const_cast<float*>(encoder),
          const_cast<float*>(event_decoder), const_cast<float*>(prong_decoder), g_position, g_combined, g_encoder,
          g_event_decoder, g_prong_decoder, prong_mask, static_cast<char*>(workspace), stream, p_drop, 0.f, seed};
  return w.backward(d_event_logits, d_prong_logits, d_event_embedding, d_prong_embedding);
}
