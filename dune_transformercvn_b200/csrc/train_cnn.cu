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
  size_t sA, sB, dmid, sums_scr, dwp;
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
    TCVN_CUDA(cudaMemsetAsync(dbl(T.sums_scr), 0, sizeof(double) * 2 * C, st));
    return colsums_into(0, X, ldx, col0, nullptr, 0, 0, nullptr, C, rows, hp, wp, dbl(T.sums_scr), C, st);
  }

  int bias_grad(const float* G, int ldg, int col0, int C, long long rows, int hp, int wp, float* dst) {
    TCVN_CUDA(cudaMemsetAsync(dbl(T.sums_scr), 0, sizeof(double) * C, st));
    TCVN_TRY(colsums_into(2, G, ldg, col0, nullptr, 0, 0, nullptr, C, rows, hp, wp, dbl(T.sums_scr), C, st));
    add_sums_kernel<<<ceil_div(C, 128), 128, 0, st>>>(dbl(T.sums_scr), C, dst);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  // dX (+)= backward of PReLU(BN(X)) given D = gradient of the activated value; parameter gradients of `bn`
  int bn_bwd(const float* X, int ldx, const float* D, int ldd, const float* fold, int C, double count, long long rows, int hp,
             int wp, float* dX, int lddx, bool accumulate, const BnArena& bn, int c0, int c0p) {
    double* s = dbl(T.sums_scr);
    TCVN_CUDA(cudaMemsetAsync(s, 0, sizeof(double) * 3 * C, st));
    TCVN_TRY(colsums_into(1, X, ldx, 0, D, ldd, 0, fold, C, rows, hp, wp, s, C, st));
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
      TCVN_TRY(colsums_into(0, blk, B.ctot, 0, nullptr, 0, 0, nullptr, B.c0p, rows, B.Hp, B.Wp, s1, B.ctot, st));
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
        TCVN_TRY(colsums_into(0, blk, B.ctot, L.kphys, nullptr, 0, 0, nullptr, g, rows, B.Hp, B.Wp, s1 + L.kphys, B.ctot, st));
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

struct T16Layer { size_t w1b, w1d, w2b, wd, fold1, fold2, mid_raw, mid_act; };
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
  size_t sA, sB, dmid, g2x, gt, dz0, sums_scr, dwp, parts, zeros, ones;
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
    T->z0 = take(N * P.Hs * P.Ws * C0 * 4);
    T->fold0 = take(5 * C0 * 4);
    size_t max_rc = 0, max_rows = 0, max_gt = 0, max_dw = (size_t)d.in_channels * 49 * C0;
    int max_c = C0 > d.out_features ? C0 : d.out_features;
    if (mid > max_c) max_c = mid;
    T->blocks.clear();
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      T16Block X;
      const size_t rows = N * B.R;
      X.blk = take(rows * B.ctot * 2);
      X.gblk = take(rows * B.ctot * 4);
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
    T->dz0 = take(N * P.Hs * P.Ws * C0 * 4);
    T->sums_scr = take(3 * (size_t)max_c * 8);
    T->dwp = take(max_dw * 4);
    T->parts = take((size_t)148 * 4 * 128 * 128 * 4 + 4 * 128 * 128 * 4 * 8);   // per-CTA partial sums of the wgrad kernel
    T->zeros = take(1024 * 4);
    T->ones = take(1024 * 4);
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

// G2x[m] = [ G[m+1] | G[m] | G[m-1] | 0 ] (4 x 32 bf16) from the fp32 gradient slice of a layer's 32 output channels,
// with the layer's dropout mask applied (same (seed, site, element) hash as the forward pass)
__global__ void g2x_kernel(const float* __restrict__ gblk, int ld, int col0, long long rows, int Hp, int Wp,
                           unsigned long long seed, unsigned long long stream_id, float p, const bf* __restrict__ blk,
                           const float* __restrict__ bstat, int ctot, bf* __restrict__ g2x) {
  // one thread = one row x 8 channels (two float4 loads, 16-byte stores).  The stored gradient lacks the deferred BN1
  // mean corrections of the later layers: g -= corrA + xhat * corrB (bstat = [mean | rstd | corrA | corrB] x ctot)
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 4) return;
  const int c = (int)(idx & 3) * 8;
  const long long m = idx >> 2;
  const unsigned rr = (unsigned)m % (unsigned)(Hp * Wp);
  const unsigned y = rr / (unsigned)Wp, x = rr - y * (unsigned)Wp;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!(y == 0 || y == (unsigned)(Hp - 1) || x == 0 || x == (unsigned)(Wp - 1))) {
    const float4 a = *reinterpret_cast<const float4*>(gblk + m * (long long)ld + col0 + c);
    const float4 b = *reinterpret_cast<const float4*>(gblk + m * (long long)ld + col0 + c + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    const uint4 xr = *reinterpret_cast<const uint4*>(blk + m * (long long)ld + col0 + c);
    const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
    const float* st = bstat + col0 + c;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xv = (i & 1) ? __uint_as_float(xw[i >> 1] & 0xffff0000u) : __uint_as_float(xw[i >> 1] << 16);
      v[i] -= st[2 * ctot + i] + (xv - st[i]) * st[ctot + i] * st[3 * ctot + i];
    }
    if (p > 0.f) {
      const float inv = 1.f / (1.f - p);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = drop_keep16(seed, stream_id, (unsigned long long)m * 32 + c + i, p) ? v[i] * inv : 0.f;
    }
  }
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  const uint4 hv = make_uint4(w[0], w[1], w[2], w[3]), zv = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(g2x + m * 128 + 32 + c) = hv;
  *reinterpret_cast<uint4*>(g2x + m * 128 + 96 + c) = zv;
  if (m > 0) *reinterpret_cast<uint4*>(g2x + (m - 1) * 128 + c) = hv;
  else *reinterpret_cast<uint4*>(g2x + 64 + c) = zv;
  if (m + 1 < rows) *reinterpret_cast<uint4*>(g2x + (m + 1) * 128 + 64 + c) = hv;
  else *reinterpret_cast<uint4*>(g2x + m * 128 + c) = zv;
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

// BN1 parameter gradients + the deferred corrections from the fused pass's reductions sums[3][C] (physical channels)
__global__ void bn1_param_corr_kernel(const double* __restrict__ sums, int C, int c_log, int c0, int c0p,
                                      const float* __restrict__ scale, double count, float* dgamma, float* dbeta, float* dalpha,
                                      float* __restrict__ corrA, float* __restrict__ corrB) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C) return;
  int c = -1;
  if (p < c0) c = p;
  else if (p >= c0p) c = p - (c0p - c0);
  if (c < 0 || c >= c_log) return;
  const double S0 = sums[p], S1 = sums[C + p], S2 = sums[2 * C + p];
  dbeta[c] += (float)S0;
  dgamma[c] += (float)S1;
  dalpha[c] += (float)S2;
  corrA[p] += scale[p] * (float)(S0 / count);
  corrB[p] += scale[p] * (float)(S1 / count);
}

// fp32 gradient columns [0, cols) -> bf16 [rows][pitch], zero padded
__global__ void to_bf16_pad_kernel(const float* __restrict__ src, int ld, int cols, long long rows, int pitch, bf* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * pitch) return;
  const int c = (int)(idx % pitch);
  const long long m = idx / pitch;
  dst[idx] = __float2bfloat16_rn(c < cols ? src[m * (long long)ld + c] : 0.f);
}

// conv2 weight gradient from the tensor-core layout dw[dy][k][dx*32 + n] -> += reference layout [n][k][dy][dx]
__global__ void unpack_conv2_grad_kernel(const float* __restrict__ dw, float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 32 * 128 * 9) return;
  const int t = idx % 9, k = (idx / 9) & 127, n = idx / (9 * 128);
  const int dy = t / 3, dx = t - dy * 3;
  dst[idx] += dw[(dy * 128 + k) * 128 + dx * 32 + n];
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
// use and kept for the life of the process; TCVN_WGRAD_STREAM=0 keeps everything on the caller's stream.
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

  float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  bf* h(size_t off) const { return reinterpret_cast<bf*>(ws + off); }
  double* dbl(size_t off) const { return reinterpret_cast<double*>(ws + off); }

  int finalize(const double* s1, const double* s2, int n_out, const BnArena& bn, int c0, int c0p, double count, float* fold) {
    FinArgs a;
    a.s1 = s1; a.s2 = s2; a.n_out = n_out; a.c_log = bn.c; a.c0 = c0; a.c0p = c0p; a.count = count;
    a.gamma = arena + bn.w; a.beta = arena + bn.b; a.alpha = arena + bn.alpha;
    a.eps = T.P.d.bn_eps; a.momentum = momentum;
    a.rm = arena + bn.rm; a.rv = arena + bn.rv;
    a.fold = fold;
    bn_finalize_map_kernel<<<ceil_div(n_out, 128), 128, 0, st>>>(a);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }

  int stats_scratch(const void* X, bool x_bf16, int ldx, int col0, int C, long long rows, int hp, int wp) {
    TCVN_CUDA(cudaMemsetAsync(dbl(T.sums_scr), 0, sizeof(double) * 2 * C, st));
    return colsums_typed(0, X, x_bf16, ldx, col0, nullptr, false, 0, 0, nullptr, 0, C, rows, hp, wp, dbl(T.sums_scr), C, st);
  }

  int bn_bwd(const void* X, bool x_bf16, int ldx, const void* D, bool d_bf16, int ldd, const float* fold, int fold_stride, int C,
             double count, long long rows, int hp, int wp, void* dX, bool o_bf16, int lddx, bool accumulate, const BnArena& bn,
             int c0, int c0p) {
    double* s = dbl(T.sums_scr);
    TCVN_CUDA(cudaMemsetAsync(s, 0, sizeof(double) * 3 * C, st));
    TCVN_TRY(colsums_typed(1, X, x_bf16, ldx, 0, D, d_bf16, ldd, 0, fold, fold_stride, C, rows, hp, wp, s, C, st));
    TCVN_TRY(bnact_bwd_apply_typed(D, d_bf16, ldd, 0, X, x_bf16, ldx, 0, fold, fold_stride, s, C, count, dX, o_bf16, lddx, 0,
                                   accumulate, rows, hp, wp, st));
    bn_param_grads_map_kernel<<<ceil_div(bn.c, 128), 128, 0, st>>>(s, C, bn.c, c0, c0p, garena + bn.w, garena + bn.b,
                                                                   garena + bn.alpha);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
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
    TCVN_TRY(tcvn_t_stem_conv(pixels, n, d.in_channels, d.height, d.width, f(T.w0), arena + P.conv0_b, C0, f(T.z0), nullptr,
                              nullptr, st));
    TCVN_TRY(stats_scratch(f(T.z0), false, C0, 0, C0, stem_rows, 0, 0));
    TCVN_TRY(finalize(dbl(T.sums_scr), dbl(T.sums_scr) + C0, C0, P.norm0, NOGAP, NOGAP, (double)stem_rows, f(T.fold0)));
    TCVN_TRY(pool_typed(0, f(T.z0), f(T.fold0), h(T.blocks[0].blk), true, n, C0, B0.H, B0.W, P.Hs, P.Ws, B0.ctot, st));
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const BlockPlan& B = P.blocks[b];
      const T16Block& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      bf* blk = h(X.blk);
      double* s1 = dbl(X.sums);
      double* s2 = s1 + B.ctot;
      TCVN_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * 2 * B.ctot, st));
      TCVN_TRY(colsums_typed(0, blk, true, B.ctot, 0, nullptr, false, 0, 0, nullptr, 0, B.c0p, rows, B.Hp, B.Wp, s1, B.ctot, st));
      for (size_t i = 0; i < B.layers.size(); ++i) {
        const LayerPlan& L = B.layers[i];
        const T16Layer& Y = X.layers[i];
        float* f1 = f(Y.fold1);
        TCVN_TRY(finalize(s1, s2, L.kpad, L.norm1, B.c0, B.c0p, count, f1));
        // raw conv1 output (+ bias): the BN2 statistics are taken from exactly the bf16 values conv2 will read
        // conv1 (+ bias), raw: the epilogue also accumulates the BN2 batch statistics from exactly the bf16 values it
        // stores; conv2's epilogue applies Dropout and accumulates the statistics of the 32 new concat channels
        double* sc2 = dbl(T.sums_scr);
        TCVN_CUDA(cudaMemsetAsync(sc2, 0, sizeof(double) * 2 * mid, st));
        TCVN_TRY(launch_gemm(true, blk, rows, B.ctot, B.ctot, h(Y.w1b), mid, L.kpad, L.kphys, f1, f1 + L.kpad, f1 + 2 * L.kpad,
                             arena + L.conv1_b, f(T.ones), h(Y.mid_raw), mid, mid, 1, B.Hp, B.Wp, st, sc2, mid));
        TCVN_TRY(finalize(sc2, sc2 + mid, mid, L.norm2, NOGAP, NOGAP, count, f(Y.fold2)));
        TCVN_TRY(bnact_fwd_typed(h(Y.mid_raw), true, mid, 0, f(Y.fold2), mid, mid, rows, B.Hp, B.Wp, h(Y.mid_act), true, mid, 0, st));
        TCVN_TRY(umma_conv2_fwd(h(Y.mid_act), rows, h(Y.w2b), arena + L.conv2_b, blk, B.ctot, L.kphys, B.Hp, B.Wp, B.W, st, p_drop,
                                seed, site * 4096 + b * 64 + i, s1 + L.kphys, B.ctot));
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
    TCVN_TRY(stats_scratch(f(T.lin), false, out, 0, out, n, 0, 0));
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

  int backward(const float* pixels, float* d_emb) {
    const CnnPlan& P = T.P;
    const tcvn_cnn_desc& d = P.d;
    const int n = T.n, C0 = d.init_features, mid = P.mid, out = d.out_features;
    const BlockPlan& last = P.blocks.back();
    const int nb = (int)P.blocks.size();
    float* dwp = f(T.dwp);
    // ---- tail (fp32)
    TCVN_TRY(tcvn_t_dropout(d_emb, out, 0, out, n, seed, site * 4096 + 4095, p_drop, st));
    TCVN_TRY(bn_bwd(f(T.lin), false, out, d_emb, false, out, f(T.fold_o), out, out, (double)n, n, 0, 0, d_emb, false, out, false,
                    P.out_norm, NOGAP, NOGAP));
    TCVN_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)last.ctot * out, st));
    TCVN_TRY(tcvn_t_wgrad(f(T.gap), last.ctot, n, last.ctot, 1, nullptr, nullptr, 0, 0, d_emb, out, 0, out, 0, 0, dwp, st));
    TCVN_TRY(unpack(dwp, 1, last.ctot, out, out, last.clog, last.c0, last.c0p, garena + P.lin_w));
    TCVN_TRY(gemm32(d_emb, out, n, out, f(T.lwt), last.ctot, nullptr, f(T.dgap), last.ctot));
    TCVN_TRY(pool_typed(3, f(T.dgap), nullptr, h(T.sA), true, n, last.ctot, last.H, last.W, 0, 0, last.ctot, st));
    TCVN_TRY(bn_bwd(h(T.blocks[nb - 1].blk), true, last.ctot, h(T.sA), true, last.ctot, f(T.fold_f), last.ctot, last.ctot,
                    (double)n * last.H * last.W, (long long)n * last.R, last.Hp, last.Wp, f(T.blocks[nb - 1].gblk), false,
                    last.ctot, false, P.final_norm, last.c0, last.c0p));
    for (int b = nb - 1; b >= 0; --b) {
      const BlockPlan& B = P.blocks[b];
      const T16Block& X = T.blocks[b];
      const long long rows = (long long)n * B.R;
      const double count = (double)n * B.H * B.W;
      bf* blk = h(X.blk);
      float* gblk = f(X.gblk);
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
        // gradient of the layer's 32 output channels (complete by now) -> bf16, dropout mask applied, 3 horizontal shifts
        g2x_kernel<<<(unsigned)ceil_div_ll(rows * 4, 256), 256, 0, st>>>(gblk, B.ctot, L.kphys, rows, B.Hp, B.Wp, seed,
                                                                        site * 4096 + b * 64 + i, p_drop, blk, bstat, B.ctot, g2x);
        TCVN_LAUNCH_CHECK();
        // conv biases: every convolution of the DenseNet feeds a train-mode BatchNorm (directly, or through the concat
        // buffer), which removes any per-channel constant, so their gradient is exactly zero; the reference's autograd
        // produces rounding noise there (tests/golden/forward_train.pt: 1e-16).  The bf16 walk leaves them at zero
        // instead of spending a reduction pass per layer on them (the fp32 parity walk computes them literally).
        // conv2 weight gradient: three vertical taps of the activated bottleneck map against G2x
        {
          const int cols[3] = {0, 0, 0}, shifts[3] = {-B.Wp, 0, B.Wp}, valid[3] = {128, 128, 128};
          if (ax) {
            TCVN_CUDA(cudaEventRecord(ax->e[0], st));        // g2x is complete
            TCVN_CUDA(cudaStreamWaitEvent(wst, ax->e[0], 0));
          }
          TCVN_TRY(umma_wgrad(h(Y.mid_act), rows, mid, mid, 3, cols, shifts, valid, nullptr, nullptr, nullptr, 0, g2x, 128, 128, 0,
                              f(T.parts), dwp, false, wst));
          unpack_conv2_grad_kernel<<<ceil_div(32 * 128 * 9, 256), 256, 0, wst>>>(dwp, garena + L.conv2_w);
          TCVN_LAUNCH_CHECK();
        }
        TCVN_TRY(umma_conv2_dgrad(g2x, h(Y.wd), rows, B.Hp, B.Wp, dmid, st));
        TCVN_TRY(bn_bwd(h(Y.mid_raw), true, mid, dmid, true, mid, f(Y.fold2), mid, mid, count, rows, B.Hp, B.Wp, dmid, true, mid,
                        false, L.norm2, NOGAP, NOGAP));
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
                              dmid, mid, mid, 0, f(T.parts), dwp, false, wst));
          TCVN_TRY(unpack_on(wst, dwp, 1, n_items * 128, 128, mid, L.cin, B.c0, B.c0p, garena + L.conv1_w));
          if (ax) {
            TCVN_CUDA(cudaEventRecord(ax->e[2], wst));
            side_pending = true;
          }
        }
        TCVN_TRY(launch_gemm(false, dmid, rows, mid, mid, h(Y.w1d), L.kphys, 128, 128, nullptr, nullptr, nullptr, f(T.zeros),
                             f(T.ones), h(T.sA), L.kphys, L.kphys, ceil_div(L.kphys, 128), B.Hp, B.Wp, st));
        // BN1 + PReLU1 backward, one pass: gblk[:, :k] += sc * g and the three reductions; the mean corrections of the
        // BatchNorm backward are deferred (bstat corrA / corrB) and applied where a channel's gradient is consumed
        {
          double* sm = dbl(T.sums_scr);
          TCVN_CUDA(cudaMemsetAsync(sm, 0, sizeof(double) * 3 * L.kphys, st));
          TCVN_TRY(bn1_bwd_fused(blk, B.ctot, h(T.sA), L.kphys, f1, L.kpad, L.kphys, gblk, B.ctot, rows, B.Hp, B.Wp, sm, st));
          bn1_param_corr_kernel<<<ceil_div(L.kphys, 128), 128, 0, st>>>(sm, L.kphys, L.norm1.c, B.c0, B.c0p, f1, count,
                                                                       garena + L.norm1.w, garena + L.norm1.b,
                                                                       garena + L.norm1.alpha, bstat + 2 * B.ctot,
                                                                       bstat + 3 * B.ctot);
          TCVN_LAUNCH_CHECK();
        }
      }
      if (side_pending) TCVN_CUDA(cudaStreamWaitEvent(st, ax->e[2], 0));   // join: parts / dwp / the arena slots are final
      // the block-input channels leave the block: make their gradient final
      TCVN_TRY(bn1_correct(gblk, B.ctot, blk, B.ctot, B.c0p, rows, B.Hp, B.Wp, bstat, bstat + B.ctot, bstat + 2 * B.ctot,
                           bstat + 3 * B.ctot, st));
      if (b > 0) {
        const BlockPlan& Pv = P.blocks[b - 1];
        const T16Block& Xp = T.blocks[b - 1];
        bf* gt = h(T.gt);
        to_bf16_pad_kernel<<<(unsigned)ceil_div_ll(rows * Xp.gt_pitch, 256), 256, 0, st>>>(gblk, B.ctot, Pv.toutp, rows, Xp.gt_pitch, gt);
        TCVN_LAUNCH_CHECK();
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
                        (long long)n * Pv.R, Pv.Hp, Pv.Wp, f(Xp.gblk), false, Pv.ctot, false, Pv.tnorm, Pv.c0, Pv.c0p));
      } else {
        const long long stem_rows = (long long)n * P.Hs * P.Ws;
        // the gradient of the 64 x 200 x 140 stem map is bf16 here (three dense passes over it: pool backward, the BN0
        // reductions, the BN0 apply); it feeds only conv0's weight gradient and the BN0 / PReLU0 parameter gradients
        bf* dz0 = h(T.dz0);
        TCVN_TRY(pool_typed(1, gblk, nullptr, dz0, true, n, C0, B.H, B.W, P.Hs, P.Ws, B.ctot, st));
        TCVN_TRY(bn_bwd(f(T.z0), false, C0, dz0, true, C0, f(T.fold0), C0, C0, (double)stem_rows, stem_rows, 0, 0, dz0, true, C0,
                        false, P.norm0, NOGAP, NOGAP));
        const int k0 = d.in_channels * 49;
        TCVN_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)k0 * C0, st));
        TCVN_TRY(stem_conv_typed(pixels, n, d.in_channels, d.height, d.width, f(T.w0), nullptr, C0, nullptr, dz0, true, dwp, st));
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
