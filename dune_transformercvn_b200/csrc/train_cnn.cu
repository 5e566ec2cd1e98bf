// DenseNet pixel-map embedding, TRAIN mode: forward with batch-statistics BatchNorm + dropout, and the
// hand-written backward.  fp32 on the CUDA-core primitives of train.cu (the parity path of training).
//
// Reference being reproduced: transformercvn/network/layers/dense_net.py:8-167 under autograd in .train():
//   every BatchNorm normalises with the statistics of the current batch (biased variance) and updates its
//   running buffers (momentum 0.1, unbiased variance); Dropout(p) follows conv2 of every bottleneck (:38) and
//   the output block (:161).
// Dataflow (same in-place concat buffers as the inference walk, cnn.cu):
//   * the per-channel (sum, sum^2) of a concat buffer are accumulated ONCE per channel when it is produced;
//     the BN1 of every later layer of the block (each with its own gamma/beta/running buffers) reuses them;
//   * a layer keeps its raw conv1 output `mid` (pre-BN2) for the backward; BN+PReLU are applied by the consumer
//     GEMM on operand load, never materialised;
//   * the gradient of a concat buffer is one buffer per block; layer i adds its input gradient to channels
//     [0, k_i) after reading its own output gradient from channels [k_i, k_i+32).
// Parameter gradients are ACCUMULATED into `grad_arena`, which has the arena's layout (reference state_dict
// order); slots of the running buffers are left untouched.
#include <map>
#include <mutex>
#include <stdlib.h>
#include <utility>

#include "kernels.h"
#include "plan.h"
#include "umma.h"

namespace tcvn {

namespace {

constexpr int NOGAP = 1 << 30;

struct TLayer { size_t w1, w1t, w2, w2t, fold1, fold2, mid; };
struct TBlock {
  size_t blk, gblk, sums, fold_t, pooled, wt, wtt, tb;
  std::vector<TLayer> layers;
};

struct TrainPlan {
  CnnPlan P;
  int n;
  size_t w0, z0, fold0, fold_f, gap, dgap, lin, fold_o, lw, lwt;
  std::vector<TBlock> blocks;
  size_t sA, sB, dmid, sums_scr, dwp, dparts;
  size_t bytes;

  static bool build(const tcvn_cnn_desc& d, int n, TrainPlan* T) {
    if (!CnnPlan::build(d, TCVN_FP32, n, &T->P)) return false;
    const CnnPlan& P = T->P;
    T->n = n;
    size_t w = 0;
    auto take = [&](size_t floats) { size_t o = w; w += (floats * 4 + 255) / 256 * 256; return o; };
    const size_t N = (size_t)(n < 1 ? 1 : n);
    const int C0 = d.init_features, mid = P.mid, g = d.growth;
    T->w0 = take((size_t)d.in_channels * 49 * C0);
    T->z0 = take(N * P.Hs * P.Ws * C0);
    T->fold0 = take(5 * C0);
    size_t max_rows_c = N * P.Hs * P.Ws * C0, max_rows_mid = 0, max_dw = (size_t)d.in_channels * 49 * C0;
    int max_c = C0 > d.out_features ? C0 : d.out_features;
    if (mid > max_c) max_c = mid;
    T->blocks.clear();
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      TBlock X;
      const size_t rows = N * B.R;
      X.blk = take(rows * B.ctot);
      X.gblk = take(rows * B.ctot);
      X.sums = take(2 * 2 * (size_t)B.ctot);  // doubles
      if (rows * B.ctot > max_rows_c) max_rows_c = rows * B.ctot;
      if (rows * mid > max_rows_mid) max_rows_mid = rows * mid;
      if (B.ctot > max_c) max_c = B.ctot;
      for (const LayerPlan& L : B.layers) {
        TLayer Y;
        Y.w1 = take((size_t)L.kphys * mid);
        Y.w1t = take((size_t)L.kphys * mid);
        Y.w2 = take((size_t)9 * mid * g);
        Y.w2t = take((size_t)9 * mid * g);
        Y.fold1 = take(5 * (size_t)L.kphys);
        Y.fold2 = take(5 * (size_t)mid);
        Y.mid = take(rows * mid);
        if ((size_t)L.kphys * mid > max_dw) max_dw = (size_t)L.kphys * mid;
        if ((size_t)9 * mid * g > max_dw) max_dw = (size_t)9 * mid * g;
        X.layers.push_back(Y);
      }
      X.fold_t = X.pooled = X.wt = X.wtt = X.tb = 0;
      if (B.has_transition) {
        const BlockPlan& Nx = P.blocks[b + 1];
        X.fold_t = take(5 * (size_t)B.ctot);
        X.pooled = take(N * Nx.R * B.ctot);
        X.wt = take((size_t)B.ctot * B.toutp);
        X.wtt = take((size_t)B.ctot * B.toutp);
        X.tb = take(B.toutp);
        if ((size_t)B.ctot * B.toutp > max_dw) max_dw = (size_t)B.ctot * B.toutp;
      }
      T->blocks.push_back(X);
    }
    const BlockPlan& last = P.blocks.back();
    T->fold_f = take(5 * (size_t)last.ctot);
    T->gap = take(N * last.ctot);
    T->dgap = take(N * last.ctot);
    T->lin = take(N * d.out_features);
    T->fold_o = take(5 * (size_t)d.out_features);
    T->lw = take((size_t)last.ctot * d.out_features);
    T->lwt = take((size_t)last.ctot * d.out_features);
    if ((size_t)last.ctot * d.out_features > max_dw) max_dw = (size_t)last.ctot * d.out_features;
    T->sA = take(max_rows_c);
    T->sB = take(max_rows_c);
    T->dmid = take(max_rows_mid);
    T->sums_scr = take(2 * 3 * (size_t)max_c);  // doubles
    T->dwp = take(max_dw);
    T->dparts = take(colsum_parts_bytes() / 4);
    T->bytes = w;
    return true;
  }
};

// ---- small kernels of the walk -------------------------------------------------------------------------------
// batch statistics -> fold [scale | shift | alpha | mean | rstd] over n_out PHYSICAL channels (alignment-padding
// channels get an all-zero fold), running-buffer update over the logical channels
struct FinArgs {
  const double* s1; const double* s2;  // sum x, sum x^2 by physical channel
  int n_out, c_log, c0, c0p;
  double count;
  const float *gamma, *beta, *alpha;
  float eps, momentum;
  float *rm, *rv;
  float* fold;
};

__global__ void bn_finalize_map_kernel(const FinArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n_out) return;
  int c = -1;
  if (p < a.c0) c = p;
  else if (p >= a.c0p) c = p - (a.c0p - a.c0);
  float sc = 0.f, sh = 0.f, al = 0.f, mu = 0.f, rs = 0.f;
  if (c >= 0 && c < a.c_log) {
    const double mean = a.s1[p] / a.count;
    double var = a.s2[p] / a.count - mean * mean;
    if (var < 0.0) var = 0.0;
    rs = (float)(1.0 / sqrt(var + (double)a.eps));
    sc = a.gamma[c] * rs;
    sh = a.beta[c] - (float)mean * sc;
    al = a.alpha[c];
    mu = (float)mean;
    if (a.rm) {
      const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
      a.rm[c] = (1.f - a.momentum) * a.rm[c] + a.momentum * (float)mean;
      a.rv[c] = (1.f - a.momentum) * a.rv[c] + a.momentum * (float)unbiased;
    }
  }
  a.fold[p] = sc;
  a.fold[a.n_out + p] = sh;
  a.fold[2 * a.n_out + p] = al;
  a.fold[3 * a.n_out + p] = mu;
  a.fold[4 * a.n_out + p] = rs;
}

// BN + PReLU parameter gradients from the backward reductions sums[3][stride] (physical channels)
__global__ void bn_param_grads_map_kernel(const double* __restrict__ sums, int stride, int c_log, int c0, int c0p,
                                          float* dgamma, float* dbeta, float* dalpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_log) return;
  const int p = c < c0 ? c : c + (c0p - c0);
  dbeta[c] += (float)sums[p];
  dgamma[c] += (float)sums[stride + p];
  dalpha[c] += (float)sums[2 * stride + p];
}

// weight gradient in kernel layout [taps][K_phys][N_phys] -> += reference layout [n_log][k_log][taps]
__global__ void unpack_grad_kernel(const float* __restrict__ src, int taps, int K_phys, int N_phys, int n_log, int k_log,
                                   int c0, int c0p, float* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_log * k_log * taps) return;
  const int t = (int)(idx % taps);
  const int k = (int)((idx / taps) % k_log);
  const int n = (int)(idx / ((long long)taps * k_log));
  const int kp = k < c0 ? k : k + (c0p - c0);
  dst[idx] += src[((size_t)t * K_phys + kp) * N_phys + n];
}

__global__ void add_sums_kernel(const double* __restrict__ sums, int c, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) dst[i] += (float)sums[i];
}

struct TWalk {
  const TrainPlan& T;
  float* arena;
  float* garena;
  char* ws;
  cudaStream_t st;
  float p_drop, momentum;
  uint64_t seed, site;

  float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  double* dbl(size_t off) const { return reinterpret_cast<double*>(ws + off); }

  int finalize(const double* s1, const double* s2, int n_out, const BnArena& bn, int c0, int c0p, double count, float* fold,
               bool update) {
    FinArgs a;
    a.s1 = s1; a.s2 = s2; a.n_out = n_out; a.c_log = bn.c; a.c0 = c0; a.c0p = c0p; a.count = count;
    a.gamma = arena + bn.w; a.beta = arena + bn.b; a.alpha = arena + bn.alpha;
    a.eps = T.P.d.bn_eps; a.momentum = momentum;
    a.rm = update ? arena + bn.rm : nullptr; a.rv = update ? arena + bn.rv : nullptr;
    a.fold = fold;
    bn_finalize_map_kernel<<<ceil_div(n_out, 128), 128, 0, st>>>(a);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  int gemm(const float* A, int lda, long long rows, int K, int taps, const int* tap_off, const float* W, int N,
           const float* a_fold, int a_hp, int a_wp, const float* bias, float* out, int ldo, int col0, int o_hp, int o_wp,
           bool accumulate = false) {
    return tcvn_t_gemm(A, lda, rows, K, taps, tap_off, W, N, a_fold, a_hp, a_wp, bias, out, ldo, col0, o_hp, o_wp,
                       accumulate ? 1 : 0, st);
  }

  int stats_scratch(const float* X, int ldx, int col0, int C, long long rows, int hp, int wp) {
    return colsums_into(0, X, ldx, col0, nullptr, 0, 0, nullptr, C, rows, hp, wp, dbl(T.sums_scr), C, dbl(T.dparts), st);
  }

  int bias_grad(const float* G, int ldg, int col0, int C, long long rows, int hp, int wp, float* dst) {
    TCVN_TRY(colsums_into(2, G, ldg, col0, nullptr, 0, 0, nullptr, C, rows, hp, wp, dbl(T.sums_scr), C, dbl(T.dparts), st));
    add_sums_kernel<<<ceil_div(C, 128), 128, 0, st>>>(dbl(T.sums_scr), C, dst);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // dX (+)= backward of PReLU(BN(X)) given D = gradient of the activated value; parameter gradients of `bn`
  int bn_bwd(const float* X, int ldx, const float* D, int ldd, const float* fold, int C, double count, long long rows, int hp,
             int wp, float* dX, int lddx, bool accumulate, const BnArena& bn, int c0, int c0p) {
    double* s = dbl(T.sums_scr);
    TCVN_TRY(colsums_into(1, X, ldx, 0, D, ldd, 0, fold, C, rows, hp, wp, s, C, dbl(T.dparts), st));
    TCVN_TRY(tcvn_t_bnact_bwd_apply(D, ldd, 0, X, ldx, 0, fold, s, C, count, dX, lddx, 0, accumulate ? 1 : 0, rows, hp, wp,
                                    nullptr, nullptr, nullptr, st));
    bn_param_grads_map_kernel<<<ceil_div(bn.c, 128), 128, 0, st>>>(s, C, bn.c, c0, c0p, garena + bn.w, garena + bn.b,
                                                                   garena + bn.alpha);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // dW (reference layout, in garena) += A^T G computed in kernel layout through the scratch buffer
  int wgrad(const float* A, int lda, long long rows, int K_phys, int taps, const int* tap_off, const float* a_fold, int a_hp,
            int a_wp, const float* G, int ldg, int gcol0, int N_phys, int g_hp, int g_wp, int n_log, int k_log, int c0,
            int c0p, float* dst) {
    float* dwp = f(T.dwp);
    TCVN_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * K_phys * N_phys, st));
    TCVN_TRY(tcvn_t_wgrad(A, lda, rows, K_phys, taps, tap_off, a_fold, a_hp, a_wp, G, ldg, gcol0, N_phys, g_hp, g_wp, dwp, st));
    const long long total = (long long)n_log * k_log * taps;
    unpack_grad_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(dwp, taps, K_phys, N_phys, n_log, k_log, c0, c0p, dst);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  int pack() {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int mid = P.mid, g = d.growth;
    TCVN_TRY(repack(arena + P.conv0_w, d.init_features, d.in_channels * 49, 1, NOGAP, NOGAP, d.in_channels * 49,
                    d.init_features, false, false, f(T.w0), st));
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      const TBlock& X = T.blocks[b];
      for (size_t i = 0; i < B.layers.size(); ++i) {
        const LayerPlan& L = B.layers[i];
        const TLayer& Y = X.layers[i];
        TCVN_TRY(repack(arena + L.conv1_w, mid, L.cin, 1, B.c0, B.c0p, L.kphys, mid, false, false, f(Y.w1), st));
        TCVN_TRY(repack(arena + L.conv1_w, mid, L.cin, 1, B.c0, B.c0p, L.kphys, mid, true, false, f(Y.w1t), st));
        TCVN_TRY(repack(arena + L.conv2_w, g, mid, 9, NOGAP, NOGAP, mid, g, false, false, f(Y.w2), st));
        TCVN_TRY(repack(arena + L.conv2_w, g, mid, 9, NOGAP, NOGAP, mid, g, true, false, f(Y.w2t), st));
      }
      if (B.has_transition) {
        TCVN_TRY(repack(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.ctot, B.toutp, false, false, f(X.wt), st));
        TCVN_TRY(repack(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.ctot, B.toutp, true, false, f(X.wtt), st));
        TCVN_TRY(pad_copy(arena + B.tconv_b, B.tout, f(X.tb), B.toutp, st));
      }
    }
    const BlockPlan& last = P.blocks.back();
    TCVN_TRY(repack(arena + P.lin_w, d.out_features, last.clog, 1, last.c0, last.c0p, last.ctot, d.out_features, false,
                    false, f(T.lw), st));
    TCVN_TRY(repack(arena + P.lin_w, d.out_features, last.clog, 1, last.c0, last.c0p, last.ctot, d.out_features, true,
                    false, f(T.lwt), st));
    return TCVN_OK;
  }

  int forward(const float* pixels, float* emb) {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int n = T.n, C0 = d.init_features, mid = P.mid, g = d.growth;
    TCVN_TRY(pack());
    // ---- stem: raw conv0 -> batch statistics -> BN0 + PReLU0 + AvgPool(3,2) into block 0
    const BlockPlan& B0 = P.blocks[0];
    const long long stem_rows = (long long)n * P.Hs * P.Ws;
    TCVN_TRY(tcvn_t_stem_conv(pixels, n, d.in_channels, d.height, d.width, f(T.w0), arena + P.conv0_b, C0, f(T.z0), nullptr,
                              nullptr, st));
    TCVN_TRY(stats_scratch(f(T.z0), C0, 0, C0, stem_rows, 0, 0));
    TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + C0, C0, P.norm0, NOGAP, NOGAP, (double)stem_rows, f(T.fold0), true));
    TCVN_TRY(tcvn_t_pool(0, f(T.z0), f(T.fold0), f(T.blocks[0].blk), n, C0, B0.H, B0.W, P.Hs, P.Ws, B0.ctot, st));
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      const TBlock& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      float* blk = f(X.blk);
      double* s1 = dbl(X.sums);
      double* s2 = s1 + B.ctot;
      TCVN_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * 2 * B.ctot, st));
      TCVN_TRY(colsums_into(0, blk, B.ctot, 0, nullptr, 0, 0, nullptr, B.c0p, rows, B.Hp, B.Wp, s1, B.ctot, dbl(T.dparts), st));
      int tap_off[9];
      for (int t = 0; t < 9; ++t) tap_off[t] = (t / 3 - 1) * B.Wp + (t % 3 - 1);
      for (size_t i = 0; i < B.layers.size(); ++i) {
        const LayerPlan& L = B.layers[i];
        const TLayer& Y = X.layers[i];
        TCVN_TRY(finalize(s1, s2, L.kphys, L.norm1, B.c0, B.c0p, count, f(Y.fold1), true));
        TCVN_TRY(gemm(blk, B.ctot, rows, L.kphys, 1, nullptr, f(Y.w1), mid, f(Y.fold1), 0, 0, arena + L.conv1_b, f(Y.mid), mid,
                      0, B.Hp, B.Wp));
        TCVN_TRY(stats_scratch(f(Y.mid), mid, 0, mid, rows, B.Hp, B.Wp));
        TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + mid, mid, L.norm2, NOGAP, NOGAP, count, f(Y.fold2), true));
        TCVN_TRY(gemm(f(Y.mid), mid, rows, mid, 9, tap_off, f(Y.w2), g, f(Y.fold2), B.Hp, B.Wp, arena + L.conv2_b, blk, B.ctot,
                      L.kphys, B.Hp, B.Wp));
        TCVN_TRY(tcvn_t_dropout(blk, B.ctot, L.kphys, g, rows, seed, site * 4096 + b * 64 + i, p_drop, st));
        TCVN_TRY(colsums_into(0, blk, B.ctot, L.kphys, nullptr, 0, 0, nullptr, g, rows, B.Hp, B.Wp, s1 + L.kphys, B.ctot, dbl(T.dparts), st));
      }
      if (B.has_transition) {
        const BlockPlan& Nx = P.blocks[b + 1];
        TCVN_TRY(finalize(s1, s2, B.ctot, B.tnorm, B.c0, B.c0p, count, f(X.fold_t), true));
        TCVN_TRY(tcvn_t_act_pool2(blk, n, B.H, B.W, B.ctot, B.ctot, f(X.fold_t), f(X.pooled), Nx.H, Nx.W, st));
        TCVN_TRY(gemm(f(X.pooled), B.ctot, (long long)n * Nx.R, B.ctot, 1, nullptr, f(X.wt), B.toutp, nullptr, 0, 0, f(X.tb),
                      f(T.blocks[b + 1].blk), Nx.ctot, 0, Nx.Hp, Nx.Wp));
      } else {
        TCVN_TRY(finalize(s1, s2, B.ctot, P.final_norm, B.c0, B.c0p, count, f(T.fold_f), true));
        TCVN_TRY(tcvn_t_act_gap(blk, n, B.H, B.W, B.ctot, B.ctot, f(T.fold_f), f(T.gap), st));
      }
    }
    // ---- tail: Linear (no bias) -> BatchNorm1d (batch statistics over the n images) -> PReLU -> Dropout
    const BlockPlan& last = P.blocks.back();
    const int out = d.out_features;
    TCVN_TRY(gemm(f(T.gap), last.ctot, n, last.ctot, 1, nullptr, f(T.lw), out, nullptr, 0, 0, nullptr, f(T.lin), out, 0, 0, 0));
    TCVN_TRY(stats_scratch(f(T.lin), out, 0, out, n, 0, 0));
    TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + out, out, P.out_norm, NOGAP, NOGAP, (double)n, f(T.fold_o), true));
    TCVN_TRY(tcvn_t_bnact_fwd(f(T.lin), out, 0, f(T.fold_o), out, n, 0, 0, emb, out, 0, st));
    TCVN_TRY(tcvn_t_dropout(emb, out, 0, out, n, seed, site * 4096 + 4095, p_drop, st));
    return TCVN_OK;
  }

  int backward(const float* pixels, float* d_emb) {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int n = T.n, C0 = d.init_features, mid = P.mid, g = d.growth, out = d.out_features;
    const BlockPlan& last = P.blocks.back();
    const int nb = (int)P.blocks.size();
    // ---- tail
    TCVN_TRY(tcvn_t_dropout(d_emb, out, 0, out, n, seed, site * 4096 + 4095, p_drop, st));
    TCVN_TRY(bn_bwd(f(T.lin), out, d_emb, out, f(T.fold_o), out, (double)n, n, 0, 0, d_emb, out, false, P.out_norm, NOGAP, NOGAP));
    TCVN_TRY(wgrad(f(T.gap), last.ctot, n, last.ctot, 1, nullptr, nullptr, 0, 0, d_emb, out, 0, out, 0, 0, out, last.clog,
                   last.c0, last.c0p, garena + P.lin_w));
    TCVN_TRY(gemm(d_emb, out, n, out, 1, nullptr, f(T.lwt), last.ctot, nullptr, 0, 0, nullptr, f(T.dgap), last.ctot, 0, 0, 0));
    TCVN_TRY(tcvn_t_pool(3, f(T.dgap), nullptr, f(T.sA), n, last.ctot, last.H, last.W, 0, 0, last.ctot, st));
    TCVN_TRY(bn_bwd(f(T.blocks[nb - 1].blk), last.ctot, f(T.sA), last.ctot, f(T.fold_f), last.ctot,
                    (double)n * last.H * last.W, (long long)n * last.R, last.Hp, last.Wp, f(T.blocks[nb - 1].gblk), last.ctot,
                    false, P.final_norm, last.c0, last.c0p));
    for (int b = nb - 1; b >= 0; --b) {
      const BlockPlan& B = P.blocks[b];
      const TBlock& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      float* blk = f(X.blk);
      float* gblk = f(X.gblk);
      float* dmid = f(T.dmid);
      int tap_off[9], tap_neg[9];
      for (int t = 0; t < 9; ++t) { tap_off[t] = (t / 3 - 1) * B.Wp + (t % 3 - 1); tap_neg[t] = -tap_off[t]; }
      for (int i = (int)B.layers.size() - 1; i >= 0; --i) {
        const LayerPlan& L = B.layers[i];
        const TLayer& Y = X.layers[i];
        // gradient of this layer's 32 output channels is complete: every later consumer has added to it
        TCVN_TRY(tcvn_t_dropout(gblk, B.ctot, L.kphys, g, rows, seed, site * 4096 + b * 64 + i, p_drop, st));
        TCVN_TRY(bias_grad(gblk, B.ctot, L.kphys, g, rows, B.Hp, B.Wp, garena + L.conv2_b));
        TCVN_TRY(wgrad(f(Y.mid), mid, rows, mid, 9, tap_off, f(Y.fold2), B.Hp, B.Wp, gblk, B.ctot, L.kphys, g, B.Hp, B.Wp, g, mid,
                       NOGAP, NOGAP, garena + L.conv2_w));
        TCVN_TRY(gemm(gblk + L.kphys, B.ctot, rows, g, 9, tap_neg, f(Y.w2t), mid, nullptr, B.Hp, B.Wp, nullptr, dmid, mid, 0, B.Hp,
                      B.Wp));
        TCVN_TRY(bn_bwd(f(Y.mid), mid, dmid, mid, f(Y.fold2), mid, count, rows, B.Hp, B.Wp, dmid, mid, false, L.norm2, NOGAP,
                        NOGAP));
        TCVN_TRY(bias_grad(dmid, mid, 0, mid, rows, B.Hp, B.Wp, garena + L.conv1_b));
        TCVN_TRY(wgrad(blk, B.ctot, rows, L.kphys, 1, nullptr, f(Y.fold1), 0, 0, dmid, mid, 0, mid, B.Hp, B.Wp, mid, L.cin, B.c0,
                       B.c0p, garena + L.conv1_w));
        TCVN_TRY(gemm(dmid, mid, rows, mid, 1, nullptr, f(Y.w1t), L.kphys, nullptr, 0, 0, nullptr, f(T.sA), L.kphys, 0, B.Hp, B.Wp));
        TCVN_TRY(bn_bwd(blk, B.ctot, f(T.sA), L.kphys, f(Y.fold1), L.kphys, count, rows, B.Hp, B.Wp, gblk, B.ctot, true, L.norm1,
                        B.c0, B.c0p));
      }
      if (b > 0) {
        // ---- transition b-1: gblk[:, :toutp] is the gradient of its 1x1 convolution output
        const BlockPlan& Pv = P.blocks[b - 1];
        const TBlock& Xp = T.blocks[b - 1];
        TCVN_TRY(bias_grad(gblk, B.ctot, 0, Pv.tout, rows, B.Hp, B.Wp, garena + Pv.tconv_b));
        TCVN_TRY(wgrad(f(Xp.pooled), Pv.ctot, rows, Pv.ctot, 1, nullptr, nullptr, 0, 0, gblk, B.ctot, 0, Pv.toutp, B.Hp, B.Wp,
                       Pv.tout, Pv.clog, Pv.c0, Pv.c0p, garena + Pv.tconv_w));
        TCVN_TRY(gemm(gblk, B.ctot, rows, Pv.toutp, 1, nullptr, f(Xp.wtt), Pv.ctot, nullptr, B.Hp, B.Wp, nullptr, f(T.sA), Pv.ctot,
                      0, B.Hp, B.Wp));
        TCVN_TRY(tcvn_t_pool(2, f(T.sA), nullptr, f(T.sB), n, Pv.ctot, Pv.H, Pv.W, B.H, B.W, Pv.ctot, st));
        TCVN_TRY(bn_bwd(f(Xp.blk), Pv.ctot, f(T.sB), Pv.ctot, f(Xp.fold_t), Pv.ctot, (double)n * Pv.H * Pv.W,
                        (long long)n * Pv.R, Pv.Hp, Pv.Wp, f(Xp.gblk), Pv.ctot, false, Pv.tnorm, Pv.c0, Pv.c0p));
      } else {
        // ---- stem: AvgPool(3,2) backward -> BN0 + PReLU0 backward -> conv0 bias / weight gradients (hit-driven)
        const long long stem_rows = (long long)n * P.Hs * P.Ws;
        TCVN_TRY(tcvn_t_pool(1, gblk, nullptr, f(T.sA), n, C0, B.H, B.W, P.Hs, P.Ws, B.ctot, st));
        TCVN_TRY(bn_bwd(f(T.z0), C0, f(T.sA), C0, f(T.fold0), C0, (double)stem_rows, stem_rows, 0, 0, f(T.sA), C0, false, P.norm0,
                        NOGAP, NOGAP));
        TCVN_TRY(bias_grad(f(T.sA), C0, 0, C0, stem_rows, 0, 0, garena + P.conv0_b));
        float* dwp = f(T.dwp);
        const int k0 = d.in_channels * 49;
        TCVN_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)k0 * C0, st));
        TCVN_TRY(tcvn_t_stem_conv(pixels, n, d.in_channels, d.height, d.width, f(T.w0), nullptr, C0, nullptr, f(T.sA), dwp, st));
        unpack_grad_kernel<<<ceil_div(k0 * C0, 256), 256, 0, st>>>(dwp, 1, k0, C0, C0, k0, NOGAP, NOGAP, garena + P.conv0_w);
        TCVN_LAUNCH_CHECK();
      }
    }
    return TCVN_OK;
  }
};


// =================================================================================================
// bf16 training walk: activations (concat buffers, bottleneck maps) in bf16, every convolution of the dense
// blocks and transitions on tcgen05 (forward: umma.cu kernels with batch-statistics folds; backward: conv2 input
// gradient + MN-major weight-gradient kernels of umma_train.cu, conv1 / transition input gradients through the
// plain K-major GEMM).  Gradients of the concat buffers stay fp32 (they are accumulated across up to 12 layers);
// statistics, folds and all parameter gradients are fp32/fp64 as in the fp32 walk.  The stem and the tail (<3 % of
// the FLOPs) run the fp32 kernels.
// =================================================================================================
typedef __nv_bfloat16 bf;

constexpr int kPullMax = 16;            // later layers of a block a gradient can be pulled from
constexpr int kPullCtasMax = 148 * 2;   // grid (and bias-gradient slots) of grad_pull_kernel
constexpr long long kFuseDgradRows = 160000;   // up to here the BN2 backward reductions are fused into the dgrad epilogue

struct T16Layer { size_t w1b, w1d, w2b, wd, fold1, fold2, mid_raw, mid_act, dA1; };
struct T16Block {
  size_t blk, gblk, sums, fold_t, pooled, wtb, wtd, tb16;
  size_t bstat;   // [mean | rstd | corrA | corrB] x ctot floats: block statistics and the deferred BN1 corrections
  int gt_pitch;
  std::vector<T16Layer> layers;
};

struct TrainPlan16 {
  CnnPlan P;
  int n;
  size_t w0, z0, fold0, fold_f, gap, dgap, lin, fold_o, lw, lwt;
  std::vector<T16Block> blocks;
  size_t sA, sB, dmid, g2x, gt, dz0, sums_scr, dwp, parts, zeros, ones, dparts, bias_parts;
  size_t bytes;

  static bool build(const tcvn_cnn_desc& d, int n, TrainPlan16* T) {
    if (!CnnPlan::build(d, TCVN_BF16, n, &T->P)) return false;
    const CnnPlan& P = T->P;
    if (P.mid != 128 || d.growth != 32) return false;   // the tcgen05 kernels are specialised for 128 / 32
    T->n = n;
    size_t w = 0;
    auto take = [&](size_t bytes) { size_t o = w; w += (bytes + 1023) / 1024 * 1024; return o; };
    const size_t N = (size_t)(n < 1 ? 1 : n);
    const int C0 = d.init_features, mid = P.mid, g = d.growth;
    T->w0 = take((size_t)d.in_channels * 49 * C0 * 4);
    T->z0 = take(N * P.Hs * P.Ws * C0 * 2);
    T->fold0 = take(5 * C0 * 4);
    size_t max_rc = 0, max_rows = 0, max_gt = 0, max_dw = (size_t)d.in_channels * 49 * C0;
    int max_c = C0 > d.out_features ? C0 : d.out_features;
    if (mid > max_c) max_c = mid;
    T->blocks.clear();
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      if ((int)B.layers.size() > kPullMax + 1) return false;
      T16Block X;
      const size_t rows = N * B.R;
      X.blk = take(rows * B.ctot * 2);
      X.gblk = take(rows * B.ctot * 2);   // gradient of the block's channels as its consumer left it (bf16; only READ by the layers)
      X.sums = take(2 * (size_t)B.ctot * 8);
      X.bstat = take(4 * (size_t)B.ctot * 4);
      if (rows * B.ctot > max_rc) max_rc = rows * B.ctot;
      if (rows > max_rows) max_rows = rows;
      if (B.ctot > max_c) max_c = B.ctot;
      for (const LayerPlan& L : B.layers) {
        T16Layer Y;
        Y.w1b = take((size_t)mid * L.kpad * 2);
        Y.w1d = take((size_t)L.kphys * mid * 2);
        Y.w2b = take((size_t)9 * g * mid * 2);
        Y.wd = take((size_t)3 * 128 * 128 * 2);
        Y.fold1 = take(5 * (size_t)L.kpad * 4);
        Y.fold2 = take(5 * (size_t)mid * 4);
        Y.mid_raw = take(rows * mid * 2);
        Y.mid_act = take(rows * mid * 2);
        Y.dA1 = take(rows * L.kphys * 2);   // gradient of the layer's activated conv1 input: pulled by the earlier layers
        X.layers.push_back(Y);
      }
      X.fold_t = X.pooled = X.wtb = X.wtd = X.tb16 = 0;
      X.gt_pitch = 0;
      if (B.has_transition) {
        const BlockPlan& Nx = P.blocks[b + 1];
        X.gt_pitch = round_up(B.toutp, 64);
        X.fold_t = take(5 * (size_t)B.ctot * 4);
        X.pooled = take(N * Nx.R * B.ctot * 2);
        X.wtb = take((size_t)B.tn_tiles * 128 * B.tkpad * 2);
        X.wtd = take((size_t)B.ctot * X.gt_pitch * 2);
        X.tb16 = take((size_t)B.tn_tiles * 128 * 4);
        if ((size_t)B.ctot * B.toutp > max_dw) max_dw = (size_t)B.ctot * B.toutp;
        if (N * Nx.R * X.gt_pitch > max_gt) max_gt = N * Nx.R * X.gt_pitch;
        if (N * Nx.R * B.ctot > max_rc) max_rc = N * Nx.R * B.ctot;
      }
      T->blocks.push_back(X);
    }
    const BlockPlan& last = P.blocks.back();
    T->fold_f = take(5 * (size_t)last.ctot * 4);
    T->gap = take(N * last.ctot * 4);
    T->dgap = take(N * last.ctot * 4);
    T->lin = take(N * d.out_features * 4);
    T->fold_o = take(5 * (size_t)d.out_features * 4);
    T->lw = take((size_t)last.ctot * d.out_features * 4);
    T->lwt = take((size_t)last.ctot * d.out_features * 4);
    if ((size_t)last.ctot * d.out_features > max_dw) max_dw = (size_t)last.ctot * d.out_features;
    max_dw += (size_t)4 * 128 * 128;   // a weight-gradient tile group behind the largest assembled matrix
    T->sA = take(max_rc * 2);
    T->sB = take(max_rc * 2);
    T->dmid = take(max_rows * mid * 2);
    T->g2x = take(max_rows * 128 * 2);
    T->gt = take((max_gt ? max_gt : 1) * 2);
    T->dz0 = take(N * P.Hs * P.Ws * C0 * 2);
    T->sums_scr = take(3 * (size_t)max_c * 8);
    T->dwp = take(max_dw * 4);
    T->parts = take((size_t)148 * 4 * 128 * 128 * 4 + 4 * 128 * 128 * 4 * 8);   // per-CTA partial sums of the wgrad kernels
    T->zeros = take(1024 * 4);
    T->ones = take(1024 * 4);
    T->dparts = take(colsum_parts_bytes());                 // per-slot partial sums (doubles) of every statistic / reduction
    T->bias_parts = take((size_t)kPullCtasMax * 32 * 8);    // per-CTA partial sums of a conv2 bias gradient
    T->bytes = w;
    return true;
  }
};

__device__ __forceinline__ bool drop_keep16(unsigned long long seed, unsigned long long stream_id, unsigned long long idx,
                                            float p) {
  unsigned long long z = seed * 0x100000001b3ull + stream_id * 0x9e3779b97f4a7c15ull + idx;
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  const uint32_t r = (uint32_t)((z ^ (z >> 31)) >> 32);
  return (r >> 8) * (1.0f / 16777216.0f) >= p;
}

__device__ __forceinline__ uint4 pack8_bf(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// Gradient of concat channels [col0, col0 + ncols) of a dense block, PULLED from where it was produced.
// All BN1s of a block normalise the same channels x with the same batch statistics, so the gradient of channel c is
//     g[m,c] = ginit[m,c] + sum_{j in later layers} sc_j,c * prelu'_j(sc_j,c x + sh_j,c) * dA_j[m,c]  -  corrA[c]  -  xhat[m,c] corrB[c]
// ginit = what the block's consumer (transition / final norm) left in the bf16 block gradient, dA_j = the bf16 input
// gradient of layer j's conv1 (kept per layer), corrA / corrB = the accumulated mean corrections of the BN backward
// (bn_param_reduce_kernel).  Round 1 PUSHED instead: every layer added sc_j g_j into the fp32 block gradient over all
// its input channels (read 4 + write 4 bytes per element and layer, the most expensive pass of the backward walk).
//   MODE 0  the 32 output channels of a layer -> Dropout mask (same (seed, site, element) hash as the forward) -> bf16
//           G2x[m] = [G[m+1] | G[m] | G[m-1] | 0] (the operand of conv2's input / weight gradient kernels) and the
//           per-CTA partial sums of conv2's bias gradient
//   MODE 2  the block-input channels -> bf16 [rows][pitch] zero-padded (operand of the transition's gradient GEMMs; for
//           block 0 the input of the stem's pooling backward)
// ------------------------------------------------------------------------------------------------------------------
struct PullSrc { const bf* dA; int ld; const float* fold; int fs; };
struct PullArgs {
  const bf* ginit; int ldg;
  const bf* blk; int ldb;
  const float* bstat; int ctot;
  int col0, ncols;
  long long rows; int Hp, Wp;
  int n_src; PullSrc src[kPullMax];
  bf* g2x; float p; unsigned long long seed, site; const unsigned long long* seed_off; double* bias_parts;   // MODE 0
  bf* out16; int pitch;                                                   // MODE 2
};

template <int MODE>
__global__ void __launch_bounds__(256, 2) grad_pull_kernel(const PullArgs a) {
  extern __shared__ float cst[];   // [n_src][3][ncols]: scale, shift, scale * slope
  __shared__ float red[MODE == 0 ? 256 : 1][8];
  const int tv = a.ncols >> 3, rpi = 256 / tv;
  for (int i = threadIdx.x; i < a.n_src * a.ncols; i += 256) {
    const int j = i / a.ncols, c = i - j * a.ncols;
    const float* f = a.src[j].fold;
    const int fs = a.src[j].fs;
    const float sc = f[a.col0 + c];
    cst[(j * 3 + 0) * a.ncols + c] = sc;
    cst[(j * 3 + 1) * a.ncols + c] = f[fs + a.col0 + c];
    cst[(j * 3 + 2) * a.ncols + c] = sc * f[2 * fs + a.col0 + c];
  }
  __syncthreads();
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = vx * 8;
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
  if (ry < rpi) {
    // deferred BN1 mean corrections: v -= corrA + xhat * corrB  =  v - K0 - x * K1  (x is a bf16 value: folding the mean
    // into K0 costs nothing; two constants per channel instead of four keep the kernel at two CTAs per SM without spills)
    float K0[8], K1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* st = a.bstat + a.col0 + c + i;
      const float mean = st[0], rstd = st[a.ctot], cA = st[2 * a.ctot], cB = st[3 * a.ctot];
      K1[i] = rstd * cB;
      K0[i] = fmaf(-mean, K1[i], cA);
    }
    const unsigned R = (unsigned)(a.Hp * a.Wp);
    const float inv_keep = 1.f / (1.f - a.p);
    const unsigned long long seed = seed_with_offset(a.seed, a.seed_off);
    // U rows in flight per thread: all their loads are issued before any is consumed (ncu, round 2: with one row per
    // iteration the kernel ran at 31-41 % of the HBM bandwidth, latency-bound at 22 % occupancy)
    constexpr int U = 2;
    const long long stride = (long long)gridDim.x * rpi;
    for (long long m0 = (long long)blockIdx.x * rpi + ry; m0 < a.rows; m0 += U * stride) {
      bool ring[U], inside[U];
      uint4 gi[U], xi[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long m = m0 + u * stride;
        inside[u] = m < a.rows;
        const unsigned rr = (unsigned)(inside[u] ? m : 0) % R;
        const unsigned y = rr / (unsigned)a.Wp, x_ = rr - y * (unsigned)a.Wp;
        ring[u] = y == 0 || y == (unsigned)(a.Hp - 1) || x_ == 0 || x_ == (unsigned)(a.Wp - 1);
        if (inside[u] && !ring[u]) {
          gi[u] = *reinterpret_cast<const uint4*>(a.ginit + m * (long long)a.ldg + a.col0 + c);
          xi[u] = *reinterpret_cast<const uint4*>(a.blk + m * (long long)a.ldb + a.col0 + c);
        }
      }
      float v[U][8], xv[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[u][i] = 0.f; xv[u][i] = 0.f; }
        if (inside[u] && !ring[u]) {
          const uint32_t gw[4] = {gi[u].x, gi[u].y, gi[u].z, gi[u].w}, xw[4] = {xi[u].x, xi[u].y, xi[u].z, xi[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[u][2 * i] = __uint_as_float(gw[i] << 16); v[u][2 * i + 1] = __uint_as_float(gw[i] & 0xffff0000u);
            xv[u][2 * i] = __uint_as_float(xw[i] << 16); xv[u][2 * i + 1] = __uint_as_float(xw[i] & 0xffff0000u);
          }
        }
      }
      // the sources one ahead: the gradient rows of source j + 1 are in flight while source j is consumed
      uint4 di[U], dn[U];
      auto fetch = [&](int j, uint4 (&d)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (inside[u] && !ring[u])
            d[u] = *reinterpret_cast<const uint4*>(a.src[j].dA + (m0 + u * stride) * (long long)a.src[j].ld + a.col0 + c);
      };
      if (a.n_src > 0) fetch(0, di);
      for (int j = 0; j < a.n_src; ++j) {
        if (j + 1 < a.n_src) fetch(j + 1, dn);
        const float* k = cst + (j * 3) * a.ncols + c;
        const float4 s0 = *reinterpret_cast<const float4*>(k), s1 = *reinterpret_cast<const float4*>(k + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(k + a.ncols), h1 = *reinterpret_cast<const float4*>(k + a.ncols + 4);
        const float4 a0 = *reinterpret_cast<const float4*>(k + 2 * a.ncols), a1 = *reinterpret_cast<const float4*>(k + 2 * a.ncols + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float sa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!(inside[u] && !ring[u])) continue;
          const uint32_t dw[4] = {di[u].x, di[u].y, di[u].z, di[u].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d = (i & 1) ? __uint_as_float(dw[i >> 1] & 0xffff0000u) : __uint_as_float(dw[i >> 1] << 16);
            v[u][i] = fmaf(d, fmaf(xv[u][i], sc[i], sh[i]) >= 0.f ? sc[i] : sa[i], v[u][i]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) di[u] = dn[u];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!inside[u]) continue;
        const long long m = m0 + u * stride;
        if (!ring[u]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] -= fmaf(xv[u][i], K1[i], K0[i]);
          if (MODE == 0) {
            if (a.p > 0.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[u][i] = drop_keep16(seed, a.site, (unsigned long long)m * 32 + c + i, a.p) ? v[u][i] * inv_keep : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) bsum[i] += v[u][i];
          }
        }
        if (MODE == 0) {
          const uint4 hv = pack8_bf(v[u]), zv = make_uint4(0u, 0u, 0u, 0u);
          bf* g2x = a.g2x;
          *reinterpret_cast<uint4*>(g2x + m * 128 + 32 + c) = hv;
          *reinterpret_cast<uint4*>(g2x + m * 128 + 96 + c) = zv;
          if (m > 0) *reinterpret_cast<uint4*>(g2x + (m - 1) * 128 + c) = hv;
          else *reinterpret_cast<uint4*>(g2x + 64 + c) = zv;
          if (m + 1 < a.rows) *reinterpret_cast<uint4*>(g2x + (m + 1) * 128 + 64 + c) = hv;
          else *reinterpret_cast<uint4*>(g2x + m * 128 + c) = zv;
        } else {
          *reinterpret_cast<uint4*>(a.out16 + m * (long long)a.pitch + c) = pack8_bf(v[u]);
          if (vx == 0)
            for (int z = a.ncols; z < a.pitch; z += 8)
              *reinterpret_cast<uint4*>(a.out16 + m * (long long)a.pitch + z) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  if (MODE == 0) {
    // conv2 bias gradient: the CTA's row lanes are added in a fixed order, one slot per CTA
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = bsum[i];
    __syncthreads();
    if (ry == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double s = 0.0;
        for (int r = 0; r < rpi; ++r) s += (double)red[r * tv + vx][i];
        a.bias_parts[(size_t)blockIdx.x * a.ncols + c + i] = s;
      }
    }
  }
}

// block statistics for the backward: bstat = [mean | rstd | corrA = 0 | corrB = 0] x ctot from the forward's (sum, sum^2)
__global__ void blk_stats_kernel(const double* __restrict__ s1, const double* __restrict__ s2, double count, float eps, int ctot,
                                 float* __restrict__ bstat) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ctot) return;
  const double mean = s1[p] / count;
  double var = s2[p] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  bstat[p] = (float)mean;
  bstat[ctot + p] = (float)(1.0 / sqrt(var + (double)eps));
  bstat[2 * ctot + p] = 0.f;
  bstat[3 * ctot + p] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
// Consumers of the per-slot partial sums.  A block owns 16 columns; its 16 slices each add every 16th slot, then the
// slice sums are added in order: the summation order is fixed, so every statistic is bit-reproducible.
// ------------------------------------------------------------------------------------------------------------------
// batch statistics -> fold over n_out physical channels + running-buffer update.  Channels [fresh.col0, +fresh.ncols) take
// their (sum, sum^2) from the producer's slots, parts[slot][2][Cp], and store them into s1 / s2; all others read s1 / s2.
struct Fresh { const double* parts; int n_slots, Cp, col0, ncols; };
struct FinArgs16 {
  double* s1; double* s2;
  Fresh fresh;
  int n_out, c_log, c0, c0p;
  double count;
  const float *gamma, *beta, *alpha;
  float eps, momentum;
  float *rm, *rv;
  float* fold;
};

constexpr int kRedSlices = 64;   // slices per 16-column block of the slot reductions (1024 threads)

__global__ void __launch_bounds__(1024) bn_finalize16_kernel(const FinArgs16 a) {
  __shared__ double red[2][kRedSlices][17];
  const int col = threadIdx.x & 15, slice = threadIdx.x >> 4;
  const int p = blockIdx.x * 16 + col;
  const bool fresh = a.fresh.parts != nullptr && p < a.n_out && p >= a.fresh.col0 && p < a.fresh.col0 + a.fresh.ncols;
  double x1 = 0.0, x2 = 0.0;
  if (fresh) {
    const int q = p - a.fresh.col0;
    const double* base = a.fresh.parts + q;
    const size_t cp = (size_t)a.fresh.Cp;
    int s = slice;
    for (; s + kRedSlices * 3 < a.fresh.n_slots; s += kRedSlices * 4) {   // 8 independent loads in flight, added in slot order
      double v[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[2 * u] = base[((size_t)(s + kRedSlices * u) * 2) * cp];
        v[2 * u + 1] = base[((size_t)(s + kRedSlices * u) * 2 + 1) * cp];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { x1 += v[2 * u]; x2 += v[2 * u + 1]; }
    }
    for (; s < a.fresh.n_slots; s += kRedSlices) {
      x1 += base[((size_t)s * 2) * cp];
      x2 += base[((size_t)s * 2 + 1) * cp];
    }
  }
  red[0][slice][col] = x1;
  red[1][slice][col] = x2;
  __syncthreads();
  if (slice != 0 || p >= a.n_out) return;
  double S1 = 0.0, S2 = 0.0;
  if (fresh) {
#pragma unroll 8
    for (int k = 0; k < kRedSlices; ++k) { S1 += red[0][k][col]; S2 += red[1][k][col]; }
    a.s1[p] = S1;
    a.s2[p] = S2;
  }
  int c = -1;
  if (p < a.c0) c = p;
  else if (p >= a.c0p) c = p - (a.c0p - a.c0);
  float sc = 0.f, sh = 0.f, al = 0.f, mu = 0.f, rs = 0.f;
  if (c >= 0 && c < a.c_log) {
    if (!fresh) { S1 = a.s1[p]; S2 = a.s2[p]; }
    const double mean = S1 / a.count;
    double var = S2 / a.count - mean * mean;
    if (var < 0.0) var = 0.0;
    rs = (float)(1.0 / sqrt(var + (double)a.eps));
    sc = a.gamma[c] * rs;
    sh = a.beta[c] - (float)mean * sc;
    al = a.alpha[c];
    mu = (float)mean;
    if (a.rm) {
      const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
      a.rm[c] = (1.f - a.momentum) * a.rm[c] + a.momentum * (float)mean;
      a.rv[c] = (1.f - a.momentum) * a.rv[c] + a.momentum * (float)unbiased;
    }
  }
  a.fold[p] = sc;
  a.fold[a.n_out + p] = sh;
  a.fold[2 * a.n_out + p] = al;
  a.fold[3 * a.n_out + p] = mu;
  a.fold[4 * a.n_out + p] = rs;
}

// BN + PReLU backward reductions parts[slot][3][C] (sum g, sum g xhat, sum dA min(y,0); physical channels) ->
//   sums_out[3][C] (optional: what the elementwise half of the backward reads), dgamma / dbeta / dalpha += (logical
//   channels), the deferred BN1 corrections corrA += sc S0 / count, corrB += sc S1 / count (optional), and - second job,
//   blocks >= main_blocks - a conv2 bias gradient: dbias[c] += sum over bias_slots of bias_parts[slot][bias_C]
struct BnParamArgs {
  const double* parts; int n_slots, C;
  double* sums_out;
  int c_log, c0, c0p;
  float *dgamma, *dbeta, *dalpha;
  const float* scale; double count; float* corrA; float* corrB;
  const double* bias_parts; int bias_slots, bias_C; float* dbias;
  int main_blocks;
};

__global__ void __launch_bounds__(1024) bn_param_reduce_kernel(const BnParamArgs a) {
  __shared__ double red[3][kRedSlices][17];
  const int col = threadIdx.x & 15, slice = threadIdx.x >> 4;
  if ((int)blockIdx.x >= a.main_blocks) {
    const int c = ((int)blockIdx.x - a.main_blocks) * 16 + col;
    double x = 0.0;
    if (c < a.bias_C) {
      int s = slice;
      for (; s + kRedSlices * 3 < a.bias_slots; s += kRedSlices * 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = a.bias_parts[(size_t)(s + kRedSlices * u) * a.bias_C + c];
#pragma unroll
        for (int u = 0; u < 4; ++u) x += v[u];
      }
      for (; s < a.bias_slots; s += kRedSlices) x += a.bias_parts[(size_t)s * a.bias_C + c];
    }
    red[0][slice][col] = x;
    __syncthreads();
    if (slice == 0 && c < a.bias_C) {
      double t = 0.0;
#pragma unroll 8
      for (int k = 0; k < kRedSlices; ++k) t += red[0][k][col];
      a.dbias[c] += (float)t;
    }
    return;
  }
  const int p = blockIdx.x * 16 + col;
  double x[3] = {0.0, 0.0, 0.0};
  if (p < a.C) {
    const double* base = a.parts + p;
    const size_t cc = (size_t)a.C;
    int s = slice;
    for (; s + kRedSlices * 3 < a.n_slots; s += kRedSlices * 4) {   // 12 independent loads in flight, added in slot order
      double v[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 3; ++j) v[u][j] = base[((size_t)(s + kRedSlices * u) * 3 + j) * cc];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 3; ++j) x[j] += v[u][j];
    }
    for (; s < a.n_slots; s += kRedSlices) {
#pragma unroll
      for (int j = 0; j < 3; ++j) x[j] += base[((size_t)s * 3 + j) * cc];
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) red[j][slice][col] = x[j];
  __syncthreads();
  if (slice != 0 || p >= a.C) return;
  double S[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll 8
    for (int k = 0; k < kRedSlices; ++k) S[j] += red[j][k][col];
  if (a.sums_out) {
#pragma unroll
    for (int j = 0; j < 3; ++j) a.sums_out[(size_t)j * a.C + p] = S[j];
  }
  int c = -1;
  if (p < a.c0) c = p;
  else if (p >= a.c0p) c = p - (a.c0p - a.c0);
  if (c < 0 || c >= a.c_log) return;
  a.dbeta[c] += (float)S[0];
  a.dgamma[c] += (float)S[1];
  a.dalpha[c] += (float)S[2];
  if (a.corrA) {
    a.corrA[p] += a.scale[p] * (float)(S[0] / a.count);
    a.corrB[p] += a.scale[p] * (float)(S[1] / a.count);
  }
}

// dst[(k0 + it*128 + r) * ld + n0 + c] = dw[it][r][c] for the valid rows / columns of a weight-gradient tile group
__global__ void scatter_wgrad_tiles_kernel(const float* __restrict__ dw, int n_items, int k0, int k_total, int n0, int n_valid,
                                           float* __restrict__ dst, int ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items * 128 * 128) return;
  const int c = idx & 127, r = (idx >> 7) & 127, it = idx >> 14;
  const int k = k0 + it * 128 + r;
  if (k < k_total && c < n_valid) dst[(size_t)k * ld + n0 + c] = dw[idx];
}

__global__ void fill_f32_kernel(float* dst, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

// Side stream of the bf16 backward walk: the weight-gradient branch of a bottleneck (wgrad GEMM -> partial-sum reduction ->
// unpack into the gradient arena, twice per layer) depends only on tensors the input-gradient chain has already produced,
// so it runs beside that chain instead of in front of it.  One (stream, 3 events) set per caller stream, created on first
// use and kept for the life of the process; TCVN_WGRAD_STREAM=0 keeps everything on the caller's stream.  Both orders are
// the same arithmetic (stream order only): the result is bit-identical either way.
struct AuxStream { cudaStream_t s = nullptr; cudaEvent_t e[3] = {nullptr, nullptr, nullptr}; };
static AuxStream* aux_for(cudaStream_t main) {
  static std::map<std::pair<int, cudaStream_t>, AuxStream> pool;
  static std::mutex mu;
  static const bool enabled = [] { const char* v = getenv("TCVN_WGRAD_STREAM"); return !(v && v[0] == '0'); }();
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> g(mu);
  auto key = std::make_pair(dev, main);
  auto it = pool.find(key);
  if (it != pool.end()) return &it->second;
  AuxStream a;
  if (cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  for (int i = 0; i < 3; ++i)
    if (cudaEventCreateWithFlags(&a.e[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return &(pool[key] = a);
}

struct TWalk16 {
  const TrainPlan16& T;
  float* arena;
  float* garena;
  char* ws;
  cudaStream_t st;
  float p_drop, momentum;
  uint64_t seed, site;
  Fresh pending{nullptr, 0, 0, 0, 0};   // statistics a producer kernel left in its slots for the next finalize

  float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  bf* h(size_t off) const { return reinterpret_cast<bf*>(ws + off); }
  double* dbl(size_t off) const { return reinterpret_cast<double*>(ws + off); }

  void produced(int n_slots, int Cp, int col0, int ncols) { pending = Fresh{dbl(T.dparts), n_slots, Cp, col0, ncols}; }

  int finalize(double* s1, double* s2, int n_out, const BnArena& bn, int c0, int c0p, double count, float* fold) {
    FinArgs16 a;
    a.s1 = s1; a.s2 = s2; a.fresh = pending; a.n_out = n_out; a.c_log = bn.c; a.c0 = c0; a.c0p = c0p; a.count = count;
    a.gamma = arena + bn.w; a.beta = arena + bn.b; a.alpha = arena + bn.alpha;
    a.eps = T.P.d.bn_eps; a.momentum = momentum;
    a.rm = arena + bn.rm; a.rv = arena + bn.rv;
    a.fold = fold;
    pending = Fresh{nullptr, 0, 0, 0, 0};
    bn_finalize16_kernel<<<ceil_div(n_out, 16), 1024, 0, st>>>(a);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // reduce the mode-1 partial sums in dparts: parameter gradients of `bn`, optionally the reduced sums (for the apply
  // kernel), the deferred BN1 corrections, and a conv2 bias gradient from bias_parts
  int param_reduce(int n_slots, int C, bool want_sums, const BnArena& bn, int c0, int c0p, const float* corr_scale, double count,
                   float* corrA, float* corrB, int bias_slots, float* dbias) {
    BnParamArgs a{};
    a.parts = dbl(T.dparts); a.n_slots = n_slots; a.C = C; a.sums_out = want_sums ? dbl(T.sums_scr) : nullptr;
    a.c_log = bn.c; a.c0 = c0; a.c0p = c0p;
    a.dgamma = garena + bn.w; a.dbeta = garena + bn.b; a.dalpha = garena + bn.alpha;
    a.scale = corr_scale; a.count = count; a.corrA = corrA; a.corrB = corrB;
    a.bias_parts = dbl(T.bias_parts); a.bias_slots = bias_slots; a.bias_C = 32; a.dbias = dbias;
    a.main_blocks = ceil_div(C, 16);
    bn_param_reduce_kernel<<<a.main_blocks + (dbias ? 2 : 0), 1024, 0, st>>>(a);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // dX = backward of PReLU(BN(X)) given D = gradient of the activated value; parameter gradients of `bn`
  int bn_bwd(const void* X, bool x_bf16, int ldx, const void* D, bool d_bf16, int ldd, const float* fold, int fold_stride, int C,
             double count, long long rows, int hp, int wp, void* dX, bool o_bf16, int lddx, const BnArena& bn, int c0, int c0p) {
    int slabs = 0;
    TCVN_TRY(colsums_parts(1, X, x_bf16, ldx, 0, D, d_bf16, ldd, 0, fold, fold_stride, C, rows, hp, wp, dbl(T.dparts), &slabs, st));
    TCVN_TRY(param_reduce(slabs, C, true, bn, c0, c0p, nullptr, count, nullptr, nullptr, 0, nullptr));
    return bnact_bwd_apply_typed(D, d_bf16, ldd, 0, X, x_bf16, ldx, 0, fold, fold_stride, dbl(T.sums_scr), C, count, dX, o_bf16, lddx,
                                 0, false, rows, hp, wp, st);
  }

  int gemm32(const float* A, int lda, long long rows, int K, const float* W, int N, const float* bias, float* out, int ldo) {
    return tcvn_t_gemm(A, lda, rows, K, 1, nullptr, W, N, nullptr, 0, 0, bias, out, ldo, 0, 0, 0, 0, st);
  }

  int pack() {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int mid = P.mid, g = d.growth;
    RepackList R;   // every weight re-layout of the step in two or three launches
    R.add(arena + P.conv0_w, d.init_features, d.in_channels * 49, 1, NOGAP, NOGAP, d.in_channels * 49, d.init_features, 0, false,
          f(T.w0));
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      const T16Block& X = T.blocks[b];
      for (size_t i = 0; i < B.layers.size(); ++i) {
        const LayerPlan& L = B.layers[i];
        const T16Layer& Y = X.layers[i];
        R.add(arena + L.conv1_w, mid, L.cin, 1, B.c0, B.c0p, L.kpad, mid, 1, true, h(Y.w1b));
        R.add(arena + L.conv1_w, mid, L.cin, 1, B.c0, B.c0p, L.kphys, mid, 0, true, h(Y.w1d));
        R.add(arena + L.conv2_w, g, mid, 9, NOGAP, NOGAP, mid, g, 1, true, h(Y.w2b));
        R.add(arena + L.conv2_w, g, mid, 3, NOGAP, NOGAP, 128, 128, 2, true, h(Y.wd));
      }
      if (B.has_transition) {
        const int n16 = B.tn_tiles * 128;
        R.add(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.tkpad, n16, 1, true, h(X.wtb));
        R.add(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.ctot, X.gt_pitch, 0, true, h(X.wtd));
        TCVN_TRY(pad_copy(arena + B.tconv_b, B.tout, f(X.tb16), n16, st));
      }
    }
    const BlockPlan& last = P.blocks.back();
    R.add(arena + P.lin_w, d.out_features, last.clog, 1, last.c0, last.c0p, last.ctot, d.out_features, 0, false, f(T.lw));
    R.add(arena + P.lin_w, d.out_features, last.clog, 1, last.c0, last.c0p, last.ctot, d.out_features, 1, false, f(T.lwt));
    TCVN_TRY(R.run(st));
    fill_f32_kernel<<<4, 256, 0, st>>>(f(T.zeros), 1024, 0.f);
    TCVN_LAUNCH_CHECK();
    fill_f32_kernel<<<4, 256, 0, st>>>(f(T.ones), 1024, 1.f);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  int forward(const float* pixels, float* emb) {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int n = T.n, C0 = d.init_features, mid = P.mid;
    TCVN_TRY(pack());
    const BlockPlan& B0 = P.blocks[0];
    const long long stem_rows = (long long)n * P.Hs * P.Ws;
    int slots = 0;
    // ---- stem: raw conv0 (+ bias) and its batch statistics in one hit-driven pass, then BN0 + PReLU0 + AvgPool(3,2)
    TCVN_TRY(stem_train_forward(pixels, n, d.in_channels, d.height, d.width, f(T.w0), arena + P.conv0_b, C0, h(T.z0), dbl(T.dparts),
                                &slots, st));
    produced(slots, C0, 0, C0);
    TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + C0, C0, P.norm0, NOGAP, NOGAP, (double)stem_rows, f(T.fold0)));
    TCVN_TRY(stem_pool16_forward(h(T.z0), f(T.fold0), h(T.blocks[0].blk), n, C0, B0.H, B0.W, P.Hs, P.Ws, B0.ctot, st));
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      const T16Block& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      bf* blk = h(X.blk);
      double* s1 = dbl(X.sums);
      double* s2 = s1 + B.ctot;
      // statistics of the block-input channels; every later channel's statistics come out of the epilogue of the conv2
      // that produces it.  Each producer leaves per-slot partial sums, the next finalize adds them in a fixed order.
      TCVN_TRY(colsums_parts(0, blk, true, B.ctot, 0, nullptr, false, 0, 0, nullptr, 0, B.c0p, rows, B.Hp, B.Wp, dbl(T.dparts), &slots, st));
      produced(slots, B.c0p, 0, B.c0p);
      for (size_t i = 0; i < B.layers.size(); ++i) {
        const LayerPlan& L = B.layers[i];
        const T16Layer& Y = X.layers[i];
        float* f1 = f(Y.fold1);
        TCVN_TRY(finalize(s1, s2, L.kpad, L.norm1, B.c0, B.c0p, count, f1));
        // conv1 (+ bias), raw: the epilogue also takes the BN2 batch statistics from exactly the bf16 values it stores;
        // conv2's epilogue applies Dropout and takes the statistics of the 32 new concat channels
        double* sc2 = dbl(T.sums_scr);
        TCVN_TRY(launch_gemm(true, blk, rows, B.ctot, B.ctot, h(Y.w1b), mid, L.kpad, L.kphys, f1, f1 + L.kpad, f1 + 2 * L.kpad,
                             arena + L.conv1_b, f(T.ones), h(Y.mid_raw), mid, mid, 1, B.Hp, B.Wp, st, dbl(T.dparts), &slots));
        produced(slots, mid, 0, mid);
        TCVN_TRY(finalize(sc2, sc2 + mid, mid, L.norm2, NOGAP, NOGAP, count, f(Y.fold2)));
        TCVN_TRY(bnact_fwd_typed(h(Y.mid_raw), true, mid, 0, f(Y.fold2), mid, mid, rows, B.Hp, B.Wp, h(Y.mid_act), true, mid, 0, st));
        TCVN_TRY(umma_conv2_fwd(h(Y.mid_act), rows, h(Y.w2b), arena + L.conv2_b, blk, B.ctot, L.kphys, B.Hp, B.Wp, B.W, st, p_drop,
                                seed, site * 4096 + b * 64 + i, dbl(T.dparts), &slots));
        produced(slots, 32, L.kphys, 32);
      }
      if (B.has_transition) {
        const BlockPlan& Nx = P.blocks[b + 1];
        float* ft = f(X.fold_t);
        TCVN_TRY(finalize(s1, s2, B.ctot, B.tnorm, B.c0, B.c0p, count, ft));
        TCVN_TRY(launch_act_pool2(blk, n, B.H, B.W, B.ctot, B.ctot, ft, ft + B.ctot, ft + 2 * B.ctot, h(X.pooled), Nx.H, Nx.W,
                                  false, st));
        TCVN_TRY(launch_gemm(false, h(X.pooled), (long long)n * Nx.R, B.ctot, B.ctot, h(X.wtb), B.tn_tiles * 128, B.tkpad, B.ctot,
                             nullptr, nullptr, nullptr, f(X.tb16), f(T.ones), h(T.blocks[b + 1].blk), Nx.ctot, Nx.ctot,
                             B.tn_tiles, Nx.Hp, Nx.Wp, st));
      } else {
        float* ff = f(T.fold_f);
        TCVN_TRY(finalize(s1, s2, B.ctot, P.final_norm, B.c0, B.c0p, count, ff));
        TCVN_TRY(launch_act_gap(blk, n, B.H, B.W, B.ctot, B.ctot, ff, ff + B.ctot, ff + 2 * B.ctot, f(T.gap), false, st));
      }
    }
    const BlockPlan& last = P.blocks.back();
    const int out = d.out_features;
    TCVN_TRY(gemm32(f(T.gap), last.ctot, n, last.ctot, f(T.lw), out, nullptr, f(T.lin), out));
    TCVN_TRY(colsums_parts(0, f(T.lin), false, out, 0, nullptr, false, 0, 0, nullptr, 0, out, n, 0, 0, dbl(T.dparts), &slots, st));
    produced(slots, out, 0, out);
    TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + out, out, P.out_norm, NOGAP, NOGAP, (double)n, f(T.fold_o)));
    TCVN_TRY(tcvn_t_bnact_fwd(f(T.lin), out, 0, f(T.fold_o), out, n, 0, 0, emb, out, 0, st));
    TCVN_TRY(tcvn_t_dropout(emb, out, 0, out, n, seed, site * 4096 + 4095, p_drop, st));
    return TCVN_OK;
  }

  int unpack(const float* dwp, int taps, int K_phys, int N_phys, int n_log, int k_log, int c0, int c0p, float* dst) {
    return unpack_on(st, dwp, taps, K_phys, N_phys, n_log, k_log, c0, c0p, dst);
  }
  int unpack_on(cudaStream_t s, const float* dwp, int taps, int K_phys, int N_phys, int n_log, int k_log, int c0, int c0p,
                float* dst) {
    const long long total = (long long)n_log * k_log * taps;
    unpack_grad_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(dwp, taps, K_phys, N_phys, n_log, k_log, c0, c0p, dst);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // gradient of channels [col0, col0 + ncols) of block b, pulled from the layers after `first_src - 1` (see grad_pull_kernel)
  int pull(int mode, int b, int first_src, int col0, int ncols, long long rows, uint64_t drop_site, int* bias_slots, bf* out16,
           int pitch) {
    const BlockPlan& B = T.P.blocks[b];
    const T16Block& X = T.blocks[b];
    PullArgs a{};
    a.ginit = h(X.gblk); a.ldg = B.ctot; a.blk = h(X.blk); a.ldb = B.ctot; a.bstat = f(X.bstat); a.ctot = B.ctot;
    a.col0 = col0; a.ncols = ncols; a.rows = rows; a.Hp = B.Hp; a.Wp = B.Wp;
    a.n_src = 0;
    for (int j = first_src; j < (int)B.layers.size(); ++j) {
      if (a.n_src >= kPullMax) return fail(TCVN_ERR_UNSUPPORTED, "dense block with more than %d layers", kPullMax + 1);
      a.src[a.n_src++] = PullSrc{h(X.layers[j].dA1), B.layers[j].kphys, f(X.layers[j].fold1), B.layers[j].kpad};
    }
    a.g2x = h(T.g2x); a.p = p_drop; a.seed = seed; a.site = drop_site; a.seed_off = seed_offset_ptr(); a.bias_parts = dbl(T.bias_parts);
    a.out16 = out16; a.pitch = pitch;
    if (ncols % 8 || ncols > 2048) return fail(TCVN_ERR_UNSUPPORTED, "grad_pull: %d channels", ncols);
    const int tv = ncols / 8, rpi = 256 / tv;
    long long want = ceil_div_ll(rows, (long long)rpi * 4);
    const int grid = (int)(want < 1 ? 1 : (want > kPullCtasMax ? kPullCtasMax : want));
    const size_t smem = (size_t)a.n_src * 3 * ncols * sizeof(float);
    if (mode == 0) {
      grad_pull_kernel<0><<<grid, 256, smem, st>>>(a);
      if (bias_slots) *bias_slots = grid;
    } else {
      grad_pull_kernel<2><<<grid, 256, smem, st>>>(a);
    }
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  int backward(const float* pixels, float* d_emb) {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int n = T.n, C0 = d.init_features, mid = P.mid, out = d.out_features;
    const BlockPlan& last = P.blocks.back();
    const int nb = (int)P.blocks.size();
    float* dwp = f(T.dwp);
    // ---- tail (fp32)
    TCVN_TRY(tcvn_t_dropout(d_emb, out, 0, out, n, seed, site * 4096 + 4095, p_drop, st));
    TCVN_TRY(bn_bwd(f(T.lin), false, out, d_emb, false, out, f(T.fold_o), out, out, (double)n, n, 0, 0, d_emb, false, out,
                    P.out_norm, NOGAP, NOGAP));
    TCVN_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)last.ctot * out, st));
    TCVN_TRY(wgrad_f32(f(T.gap), last.ctot, n, last.ctot, d_emb, out, out, dwp, f(T.dparts), colsum_parts_bytes() / 4, st));
    TCVN_TRY(unpack(dwp, 1, last.ctot, out, out, last.clog, last.c0, last.c0p, garena + P.lin_w));
    TCVN_TRY(gemm32(d_emb, out, n, out, f(T.lwt), last.ctot, nullptr, f(T.dgap), last.ctot));
    TCVN_TRY(pool_typed(3, f(T.dgap), nullptr, h(T.sA), true, n, last.ctot, last.H, last.W, 0, 0, last.ctot, st));
    TCVN_TRY(bn_bwd(h(T.blocks[nb - 1].blk), true, last.ctot, h(T.sA), true, last.ctot, f(T.fold_f), last.ctot, last.ctot,
                    (double)n * last.H * last.W, (long long)n * last.R, last.Hp, last.Wp, h(T.blocks[nb - 1].gblk), true,
                    last.ctot, P.final_norm, last.c0, last.c0p));
    for (int b = nb - 1; b >= 0; --b) {
      const BlockPlan& B = P.blocks[b];
      const T16Block& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      bf* blk = h(X.blk);
      bf* dmid = h(T.dmid);
      bf* g2x = h(T.g2x);
      float* bstat = f(X.bstat);
      blk_stats_kernel<<<ceil_div(B.ctot, 128), 128, 0, st>>>(dbl(X.sums), dbl(X.sums) + B.ctot, count, d.bn_eps, B.ctot, bstat);
      TCVN_LAUNCH_CHECK();
      AuxStream* ax = aux_for(st);
      const cudaStream_t wst = ax ? ax->s : st;   // stream of the weight-gradient branch
      bool side_pending = false;
      for (int i = (int)B.layers.size() - 1; i >= 0; --i) {
        const LayerPlan& L = B.layers[i];
        const T16Layer& Y = X.layers[i];
        float* f1 = f(Y.fold1);
        if (side_pending) {   // the previous layer's weight gradients still read g2x / dmid, which this layer overwrites
          TCVN_CUDA(cudaStreamWaitEvent(st, ax->e[2], 0));
          side_pending = false;
        }
        // gradient of the layer's 32 output channels (every later layer has produced its dA by now) -> bf16, dropout mask
        // applied, 3 horizontal shifts; the same pass reduces conv2's bias gradient.  (conv1 / transition / conv0 biases
        // feed a train-mode BatchNorm directly, which removes any per-channel constant: their gradient is exactly zero and
        // stays zero here.  conv2's bias passes through Dropout first - mask * b / (1 - p) is not constant - so it has one.)
        int bias_slots = 0;
        TCVN_TRY(pull(0, b, i + 1, L.kphys, 32, rows, site * 4096 + b * 64 + i, &bias_slots, nullptr, 0));
        // conv2 weight gradient: three vertical taps of the activated bottleneck map against G2x
        {
          const int cols[3] = {0, 0, 0}, shifts[3] = {-B.Wp, 0, B.Wp}, valid[3] = {128, 128, 128};
          if (ax) {
            TCVN_CUDA(cudaEventRecord(ax->e[0], st));        // g2x is complete
            TCVN_CUDA(cudaStreamWaitEvent(wst, ax->e[0], 0));
          }
          TCVN_TRY(umma_wgrad(h(Y.mid_act), rows, mid, mid, 3, cols, shifts, valid, nullptr, nullptr, nullptr, 0, g2x, 128, 128, 0,
                              f(T.parts), garena + L.conv2_w, true, wst, 2));
        }
        // conv2 input gradient, then the BN2 + PReLU2 backward reductions over (dmid, mid_raw).  On small maps the reductions
        // ride in the dgrad epilogue (one launch and one pass over dmid less); on large maps they stay a separate
        // HBM-bound pass: ncu (profiles/r2_dgrad_fused_block1.txt) shows the fused epilogue bound by the LSU pipe - 372 warp
        // shuffles + 256 shared-memory broadcasts per row - at 0.31 ns per row against 0.25 ns for dgrad + column sums
        {
          int slabs = 0;
          if (rows <= kFuseDgradRows) {
            TCVN_TRY(umma_conv2_dgrad(g2x, h(Y.wd), rows, B.Hp, B.Wp, dmid, st, h(Y.mid_raw), f(Y.fold2), dbl(T.dparts), &slabs));
          } else {
            TCVN_TRY(umma_conv2_dgrad(g2x, h(Y.wd), rows, B.Hp, B.Wp, dmid, st));
            TCVN_TRY(colsums_parts(1, h(Y.mid_raw), true, mid, 0, dmid, true, mid, 0, f(Y.fold2), mid, mid, rows, B.Hp, B.Wp,
                                   dbl(T.dparts), &slabs, st));
          }
          // reduce the slots -> (parameter gradients, conv2 bias gradient), then the elementwise half in place
          TCVN_TRY(param_reduce(slabs, mid, true, L.norm2, NOGAP, NOGAP, nullptr, count, nullptr, nullptr, bias_slots,
                                garena + L.conv2_b));
          TCVN_TRY(bnact_bwd_apply_typed(dmid, true, mid, 0, h(Y.mid_raw), true, mid, 0, f(Y.fold2), mid, dbl(T.sums_scr), mid, count,
                                         dmid, true, mid, 0, false, rows, B.Hp, B.Wp, st));
        }
        // conv1 weight gradient: 128-channel column blocks of the concat buffer, BN1 + PReLU1 applied in SMEM
        {
          const int n_items = ceil_div(L.kphys, 128);
          int cols[4], shifts[4], valid[4];
          for (int j = 0; j < n_items; ++j) { cols[j] = 128 * j; shifts[j] = 0; valid[j] = L.kphys - 128 * j < 128 ? L.kphys - 128 * j : 128; }
          if (ax) {
            TCVN_CUDA(cudaEventRecord(ax->e[1], st));        // dmid holds the BN2 input gradient
            TCVN_CUDA(cudaStreamWaitEvent(wst, ax->e[1], 0));
          }
          TCVN_TRY(umma_wgrad(blk, rows, B.ctot, B.ctot, n_items, cols, shifts, valid, f1, f1 + L.kpad, f1 + 2 * L.kpad, L.kphys,
                              dmid, mid, mid, 0, f(T.parts), garena + L.conv1_w, true, wst, 1, mid, L.cin, B.c0, B.c0p));
          if (ax) {
            TCVN_CUDA(cudaEventRecord(ax->e[2], wst));
            side_pending = true;
          }
        }
        // gradient of the activated conv1 input, kept per layer: the earlier layers pull their share from it
        TCVN_TRY(launch_gemm(false, dmid, rows, mid, mid, h(Y.w1d), L.kphys, 128, 128, nullptr, nullptr, nullptr, f(T.zeros),
                             f(T.ones), h(Y.dA1), L.kphys, L.kphys, ceil_div(L.kphys, 128), B.Hp, B.Wp, st));
        // BN1 + PReLU1 backward, reduction half only: parameter gradients + the deferred mean corrections (bstat corrA / corrB)
        {
          int slabs = 0;
          TCVN_TRY(bn1_bwd_reduce(blk, B.ctot, h(Y.dA1), L.kphys, f1, L.kpad, L.kphys, rows, B.Hp, B.Wp, dbl(T.dparts), &slabs, st));
          TCVN_TRY(param_reduce(slabs, L.kphys, false, L.norm1, B.c0, B.c0p, f1, count, bstat + 2 * B.ctot, bstat + 3 * B.ctot, 0,
                                nullptr));
        }
      }
      if (side_pending) TCVN_CUDA(cudaStreamWaitEvent(st, ax->e[2], 0));   // join: parts / dwp / the arena slots are final
      if (b > 0) {
        const BlockPlan& Pv = P.blocks[b - 1];
        const T16Block& Xp = T.blocks[b - 1];
        bf* gt = h(T.gt);
        // the block-input channels leave the block: their final gradient, bf16 and zero-padded, for the transition's GEMMs
        TCVN_TRY(pull(2, b, 0, 0, B.c0p, rows, 0, nullptr, gt, Xp.gt_pitch));
        // transition weight gradient dW[k][n] = sum_m pooled[m, k] * Gt[m, n] on the MN-major tensor-core kernel: groups
        // of <= 4 x 128 input channels against 128-column tiles of Gt, scattered into the [ctot][toutp] scratch
        {
          float* tile = dwp + (size_t)Pv.ctot * Pv.toutp;   // [4][128][128] behind the assembled matrix
          for (int k0 = 0; k0 < Pv.ctot; k0 += 512) {
            const int n_items = ceil_div((Pv.ctot - k0 < 512 ? Pv.ctot - k0 : 512), 128);
            int cols[4], shifts[4], valid[4];
            for (int j = 0; j < n_items; ++j) {
              cols[j] = k0 + 128 * j; shifts[j] = 0;
              valid[j] = Pv.ctot - cols[j] < 128 ? Pv.ctot - cols[j] : 128;
            }
            for (int n0 = 0; n0 < Pv.toutp; n0 += 128) {
              TCVN_TRY(umma_wgrad(h(Xp.pooled), rows, Pv.ctot, Pv.ctot, n_items, cols, shifts, valid, nullptr, nullptr, nullptr, 0, gt,
                                  Xp.gt_pitch, Xp.gt_pitch, n0, f(T.parts), tile, false, st));
              scatter_wgrad_tiles_kernel<<<ceil_div(n_items * 128 * 128, 256), 256, 0, st>>>(
                  tile, n_items, k0, Pv.ctot, n0, Pv.toutp - n0 < 128 ? Pv.toutp - n0 : 128, dwp, Pv.toutp);
              TCVN_LAUNCH_CHECK();
            }
          }
        }
        TCVN_TRY(unpack(dwp, 1, Pv.ctot, Pv.toutp, Pv.tout, Pv.clog, Pv.c0, Pv.c0p, garena + Pv.tconv_w));
        TCVN_TRY(launch_gemm(false, gt, rows, Xp.gt_pitch, Xp.gt_pitch, h(Xp.wtd), Pv.ctot, Xp.gt_pitch, Xp.gt_pitch, nullptr, nullptr,
                             nullptr, f(T.zeros), f(T.ones), h(T.sA), Pv.ctot, Pv.ctot, ceil_div(Pv.ctot, 128), B.Hp, B.Wp, st));
        TCVN_TRY(pool_typed(2, h(T.sA), nullptr, h(T.sB), true, n, Pv.ctot, Pv.H, Pv.W, B.H, B.W, Pv.ctot, st));
        TCVN_TRY(bn_bwd(h(Xp.blk), true, Pv.ctot, h(T.sB), true, Pv.ctot, f(Xp.fold_t), Pv.ctot, Pv.ctot, (double)n * Pv.H * Pv.W,
                        (long long)n * Pv.R, Pv.Hp, Pv.Wp, h(Xp.gblk), true, Pv.ctot, Pv.tnorm, Pv.c0, Pv.c0p));
      } else {
        const long long stem_rows = (long long)n * P.Hs * P.Ws;
        // the stem side: final gradient of the 64 pooled stem channels (bf16), AvgPool(3,2) backward into the gradient of the
        // 64 x 200 x 140 stem map (it feeds only conv0's weight gradient and BN0 / PReLU0), BN0 + PReLU0 backward in place
        bf* g0 = h(T.sA);
        TCVN_TRY(pull(2, b, 0, 0, B.c0p, rows, 0, nullptr, g0, B.c0p));
        bf* dz0 = h(T.dz0);
        TCVN_TRY(stem_pool16_backward(g0, B.c0p, dz0, n, C0, B.H, B.W, P.Hs, P.Ws, st));
        TCVN_TRY(bn_bwd(h(T.z0), true, C0, dz0, true, C0, f(T.fold0), C0, C0, (double)stem_rows, stem_rows, 0, 0, dz0, true, C0,
                        P.norm0, NOGAP, NOGAP));
        // conv0 weight gradient, hit-driven and bit-reproducible (train_stem.cu): per-CTA partial sums, added in order
        const int k0 = d.in_channels * 49;
        int slots = 0;
        TCVN_TRY(stem_train_wgrad(pixels, n, d.in_channels, d.height, d.width, dz0, C0, f(T.parts), &slots, st));
        TCVN_TRY(reduce_parts(f(T.parts), slots, (long long)k0 * C0, dwp, false, st));
        TCVN_TRY(unpack(dwp, 1, k0, C0, C0, k0, NOGAP, NOGAP, garena + P.conv0_w));
      }
    }
    return TCVN_OK;
  }
};

}  // namespace
}  // namespace tcvn

using namespace tcvn;

extern "C" size_t tcvn_cnn_train_workspace_bytes(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images) {
  if (!d || n_images < 0) { set_error("cnn_train: bad descriptor"); return 0; }
  if (prec == TCVN_BF16) {
    TrainPlan16 T;
    if (!TrainPlan16::build(*d, n_images, &T)) { set_error("cnn_train: descriptor not supported by the bf16 training path"); return 0; }
    return T.bytes;
  }
  TrainPlan T;
  if (prec != TCVN_FP32 || !TrainPlan::build(*d, n_images, &T)) { set_error("cnn_train: bad descriptor"); return 0; }
  return T.bytes;
}

extern "C" int tcvn_cnn_train_forward(const tcvn_cnn_desc* d, tcvn_precision prec, float* arena, const float* pixels,
                                      int n_images, float p_drop, float momentum, uint64_t seed, uint64_t site,
                                      float* embedding, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && arena && pixels && embedding && workspace, "cnn_train_forward: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_FP32 || prec == TCVN_BF16, "cnn_train_forward: unknown precision");
  TCVN_CHECK_ARG(n_images >= 2, "cnn_train_forward: BatchNorm in train mode needs at least 2 images (got %d)", n_images);
  TCVN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "cnn_train_forward: dropout probability out of range");
  if (prec == TCVN_BF16) {
    TrainPlan16 T;
    TCVN_CHECK_ARG(TrainPlan16::build(*d, n_images, &T), "cnn_train_forward: descriptor not supported by the bf16 path");
    if (workspace_bytes < T.bytes)
      return fail(TCVN_ERR_WORKSPACE, "cnn_train_forward: workspace %zu < %zu bytes", workspace_bytes, T.bytes);
    TWalk16 w{T, arena, nullptr, static_cast<char*>(workspace), stream, p_drop, momentum, seed, site};
    return w.forward(pixels, embedding);
  }
  TrainPlan T;
  TCVN_CHECK_ARG(TrainPlan::build(*d, n_images, &T), "cnn_train_forward: bad descriptor");
  if (workspace_bytes < T.bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_train_forward: workspace %zu < %zu bytes", workspace_bytes, T.bytes);
  TWalk w{T, arena, nullptr, static_cast<char*>(workspace), stream, p_drop, momentum, seed, site};
  return w.forward(pixels, embedding);
}

extern "C" int tcvn_cnn_train_backward(const tcvn_cnn_desc* d, tcvn_precision prec, const float* arena, float* grad_arena,
                                       const float* pixels, int n_images, float p_drop, uint64_t seed, uint64_t site,
                                       float* d_embedding, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && arena && grad_arena && pixels && d_embedding && workspace, "cnn_train_backward: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_FP32 || prec == TCVN_BF16, "cnn_train_backward: unknown precision");
  TCVN_CHECK_ARG(n_images >= 2, "cnn_train_backward: needs the state of a train-mode forward over >= 2 images");
  if (prec == TCVN_BF16) {
    TrainPlan16 T;
    TCVN_CHECK_ARG(TrainPlan16::build(*d, n_images, &T), "cnn_train_backward: descriptor not supported by the bf16 path");
    if (workspace_bytes < T.bytes)
      return fail(TCVN_ERR_WORKSPACE, "cnn_train_backward: workspace %zu < %zu bytes", workspace_bytes, T.bytes);
    TWalk16 w{T, const_cast<float*>(arena), grad_arena, static_cast<char*>(workspace), stream, p_drop, 0.f, seed, site};
    return w.backward(pixels, d_embedding);
  }
  TrainPlan T;
  TCVN_CHECK_ARG(TrainPlan::build(*d, n_images, &T), "cnn_train_backward: bad descriptor");
  if (workspace_bytes < T.bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_train_backward: workspace %zu < %zu bytes", workspace_bytes, T.bytes);
  TWalk w{T, const_cast<float*>(arena), grad_arena, static_cast<char*>(workspace), stream, p_drop, 0.f, seed, site};
  return w.backward(pixels, d_embedding);
}
