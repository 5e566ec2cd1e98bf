// Parameter packing: reference-ordered fp32 arena -> kernel-ready block.
//   * eval-mode BatchNorm2d/1d (torch defaults; dense_net.py:19,30,85,119,147,159) + PReLU fold into
//     per-channel (scale, shift, alpha) over the PHYSICAL channels of the in-place concat buffers;
//     a conv bias in front of the BN is folded into the shift.
//   * conv / linear weights re-laid K-major (bf16 tcgen05 path) or N-major (fp32 CUDA-core path),
//     with zero rows/columns for the alignment padding channels.
#include "kernels.h"
#include "plan.h"

#include <string.h>

namespace tcvn {

struct FoldArgs {
  const float* w; const float* b; const float* rm; const float* rv; const float* alpha;
  const float* conv_bias;  // nullable; bias of the convolution feeding this BN
  int c_log;               // logical channels
  int c0, c0p;             // physical index = c (c < c0) | c + (c0p - c0)
  int n_out;               // entries to write (>= physical count; the rest is zero)
  float eps;
  float* scale; float* shift; float* alpha_out;
};

__global__ void fold_bn_kernel(const FoldArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n_out) return;
  int c = -1;
  if (p < a.c0) c = p;
  else if (p >= a.c0p) c = p - (a.c0p - a.c0);
  float s = 0.f, t = 0.f, al = 0.f;
  if (c >= 0 && c < a.c_log) {
    s = a.w[c] / sqrtf(a.rv[c] + a.eps);
    t = a.b[c] - a.rm[c] * s;
    if (a.conv_bias) t = fmaf(s, a.conv_bias[c], t);
    al = a.alpha[c];
  }
  a.scale[p] = s;
  a.shift[p] = t;
  a.alpha_out[p] = al;
}

struct RepackArgs {
  const float* src;  // [n_log][k_log][taps]
  int n_log, k_log, taps;
  int c0, c0p;       // physical k = k (k < c0) | k + (c0p - c0)
  int K_out, N_out;  // padded extents of the destination
  int k_major;       // 1: dst[tap][n][k]  (K contiguous) ; 0: dst[tap][k][n]
  int bf16;
  void* dst;
  // optional per-output-row scale = bn_w[n] / sqrt(bn_rv[n] + eps): folds the BatchNorm that FOLLOWS this conv
  const float* row_bn_w; const float* row_bn_rv; float eps;
};

__global__ void repack_kernel(const RepackArgs a) {
  const long long total = (long long)a.taps * a.K_out * a.N_out;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int tap, n, kp;
  if (a.k_major) {
    kp = (int)(idx % a.K_out);
    n = (int)((idx / a.K_out) % a.N_out);
    tap = (int)(idx / ((long long)a.K_out * a.N_out));
  } else {
    n = (int)(idx % a.N_out);
    kp = (int)((idx / a.N_out) % a.K_out);
    tap = (int)(idx / ((long long)a.K_out * a.N_out));
  }
  int k = -1;
  if (kp < a.c0) k = kp;
  else if (kp >= a.c0p) k = kp - (a.c0p - a.c0);
  float v = 0.f;
  if (k >= 0 && k < a.k_log && n < a.n_log) {
    v = a.src[((size_t)n * a.k_log + k) * a.taps + tap];
    if (a.row_bn_w) v *= a.row_bn_w[n] / sqrtf(a.row_bn_rv[n] + a.eps);
  }
  if (a.bf16) static_cast<__nv_bfloat16*>(a.dst)[idx] = __float2bfloat16_rn(v);
  else static_cast<float*>(a.dst)[idx] = v;
}

// many re-layouts in ONE launch (blockIdx.y = entry): the training walks rebuild ~120 weight layouts per CNN per step
constexpr int kRepackBatch = 96;
struct RepackBatchArgs { RepackArgs e[kRepackBatch]; };

__device__ __forceinline__ void repack_one(const RepackArgs& a, long long idx) {
  const long long total = (long long)a.taps * a.K_out * a.N_out;
  if (idx >= total) return;
  int tap, n, kp;
  if (a.k_major == 2) {
    // conv2 weight [n = 32][c = 128][dy][dx] -> Wd[dy][c][dx * 32 + n] (columns 96.. zero): input-gradient operand
    const int col = (int)(idx & 127), c = (int)((idx >> 7) & 127), dy = (int)(idx >> 14);
    float v = 0.f;
    if (col < 96) v = a.src[((size_t)((col & 31) * 128 + c) * 3 + dy) * 3 + (col >> 5)];
    static_cast<__nv_bfloat16*>(a.dst)[idx] = __float2bfloat16_rn(v);
    return;
  }
  if (a.k_major) {
    kp = (int)(idx % a.K_out);
    n = (int)((idx / a.K_out) % a.N_out);
    tap = (int)(idx / ((long long)a.K_out * a.N_out));
  } else {
    n = (int)(idx % a.N_out);
    kp = (int)((idx / a.N_out) % a.K_out);
    tap = (int)(idx / ((long long)a.K_out * a.N_out));
  }
  int k = -1;
  if (kp < a.c0) k = kp;
  else if (kp >= a.c0p) k = kp - (a.c0p - a.c0);
  float v = 0.f;
  if (k >= 0 && k < a.k_log && n < a.n_log) {
    v = a.src[((size_t)n * a.k_log + k) * a.taps + tap];
    if (a.row_bn_w) v *= a.row_bn_w[n] / sqrtf(a.row_bn_rv[n] + a.eps);
  }
  if (a.bf16) static_cast<__nv_bfloat16*>(a.dst)[idx] = __float2bfloat16_rn(v);
  else static_cast<float*>(a.dst)[idx] = v;
}

__global__ void repack_batch_kernel(const __grid_constant__ RepackBatchArgs b) {
  const RepackArgs& a = b.e[blockIdx.y];
  const long long total = (long long)a.taps * a.K_out * a.N_out;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    repack_one(a, idx);
}

__global__ void pad_copy_kernel(const float* src, int n_log, float* dst, int n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out) dst[i] = i < n_log ? src[i] : 0.f;
}

__global__ void fill_kernel(float* dst, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

int pad_copy(const float* src, int n_log, float* dst, int n_out, cudaStream_t st) {
  pad_copy_kernel<<<ceil_div(n_out, 128), 128, 0, st>>>(src, n_log, dst, n_out);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int fold(const float* arena, const BnArena& bn, const float* conv_bias, int c0, int c0p, int n_out, float eps,
                char* packed, size_t o_scale, size_t o_shift, size_t o_alpha, cudaStream_t st) {
  FoldArgs f;
  f.w = arena + bn.w; f.b = arena + bn.b; f.rm = arena + bn.rm; f.rv = arena + bn.rv; f.alpha = arena + bn.alpha;
  f.conv_bias = conv_bias; f.c_log = bn.c; f.c0 = c0; f.c0p = c0p; f.n_out = n_out; f.eps = eps;
  f.scale = reinterpret_cast<float*>(packed + o_scale);
  f.shift = reinterpret_cast<float*>(packed + o_shift);
  f.alpha_out = reinterpret_cast<float*>(packed + o_alpha);
  fold_bn_kernel<<<ceil_div(n_out, 128), 128, 0, st>>>(f);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int repack(const float* src, int n_log, int k_log, int taps, int c0, int c0p, int K_out, int N_out, bool k_major,
           bool bf16, void* dst, cudaStream_t st, const float* row_bn_w, const float* row_bn_rv, float eps) {
  RepackArgs r;
  r.row_bn_w = row_bn_w; r.row_bn_rv = row_bn_rv; r.eps = eps;
  r.src = src; r.n_log = n_log; r.k_log = k_log; r.taps = taps; r.c0 = c0; r.c0p = c0p; r.K_out = K_out; r.N_out = N_out;
  r.k_major = k_major; r.bf16 = bf16; r.dst = dst;
  const long long total = (long long)taps * K_out * N_out;
  repack_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(r);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

void RepackList::add(const float* src, int n_log, int k_log, int taps, int c0, int c0p, int K_out, int N_out, int k_major,
                     bool bf16, void* dst) {
  RepackArgs r;
  r.row_bn_w = nullptr; r.row_bn_rv = nullptr; r.eps = 0.f;
  r.src = src; r.n_log = n_log; r.k_log = k_log; r.taps = taps; r.c0 = c0; r.c0p = c0p; r.K_out = K_out; r.N_out = N_out;
  r.k_major = k_major; r.bf16 = bf16; r.dst = dst;
  static_assert(sizeof(RepackArgs) == sizeof(RepackList::Entry), "RepackList::Entry must mirror RepackArgs");
  Entry e;
  memcpy(&e, &r, sizeof(r));
  items.push_back(e);
}

int RepackList::run(cudaStream_t st) {
  for (size_t i0 = 0; i0 < items.size(); i0 += kRepackBatch) {
    const int n = (int)(items.size() - i0 < (size_t)kRepackBatch ? items.size() - i0 : kRepackBatch);
    RepackBatchArgs b;
    memcpy(b.e, items.data() + i0, sizeof(RepackArgs) * n);
    dim3 grid(64, n);
    repack_batch_kernel<<<grid, 256, 0, st>>>(b);
    TCVN_LAUNCH_CHECK();
  }
  items.clear();
  return TCVN_OK;
}

}  // namespace tcvn

using namespace tcvn;

extern "C" int64_t tcvn_cnn_arena_floats(const tcvn_cnn_desc* d) {
  CnnPlan P;
  if (!d || !CnnPlan::build(*d, TCVN_FP32, 1, &P)) { set_error("cnn: bad descriptor"); return -1; }
  return P.arena_floats;
}

extern "C" size_t tcvn_cnn_packed_bytes(const tcvn_cnn_desc* d, tcvn_precision prec) {
  CnnPlan P;
  if (!d || !CnnPlan::build(*d, prec, 1, &P)) { set_error("cnn: bad descriptor"); return 0; }
  return P.packed_bytes;
}

extern "C" size_t tcvn_cnn_workspace_bytes(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images) {
  CnnPlan P;
  if (!d || !CnnPlan::build(*d, prec, n_images, &P, true)) { set_error("cnn: bad descriptor"); return 0; }
  return P.ws_bytes;
}



extern "C" int tcvn_cnn_pack(const tcvn_cnn_desc* d, tcvn_precision prec, const float* arena, void* packed_v,
                             size_t packed_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && arena && packed_v, "cnn_pack: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_FP32 || prec == TCVN_BF16, "cnn_pack: unknown precision");
  CnnPlan P;
  TCVN_CHECK_ARG(CnnPlan::build(*d, prec, 1, &P), "cnn_pack: bad descriptor");
  if (packed_bytes < P.packed_bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_pack: packed buffer %zu < %zu bytes", packed_bytes, P.packed_bytes);
  char* pk = static_cast<char*>(packed_v);
  cudaStream_t st = stream;
  const bool bf = prec == TCVN_BF16;
  const float eps = d->bn_eps;
  const int NOGAP = 1 << 30;
  // stem: w0 [cin*49][C0] fp32 (tap index (c,ky,kx) == the arena's trailing dims), bias folded into BN0
  TCVN_TRY(repack(arena + P.conv0_w, d->init_features, d->in_channels * 49, 1, NOGAP, NOGAP, d->in_channels * 49,
                  d->init_features, false, false, pk + P.p_w0, st));
  TCVN_TRY(fold(arena, P.norm0, arena + P.conv0_b, NOGAP, NOGAP, d->init_features, eps, pk, P.p_s_scale, P.p_s_shift,
                P.p_s_alpha, st));
  for (auto& B : P.blocks) {
    for (auto& L : B.layers) {
      TCVN_TRY(fold(arena, L.norm1, nullptr, B.c0, B.c0p, L.kpad, eps, pk, L.p_a_scale, L.p_a_shift, L.p_a_alpha, st));
      if (bf) TCVN_TRY(repack(arena + L.conv1_w, P.mid, L.cin, 1, B.c0, B.c0p, L.kpad, P.mid, true, true, pk + L.p_w1, st,
                              arena + L.norm2.w, arena + L.norm2.rv, eps));
      else TCVN_TRY(repack(arena + L.conv1_w, P.mid, L.cin, 1, B.c0, B.c0p, L.kphys, P.mid, false, false, pk + L.p_w1, st));
      TCVN_TRY(fold(arena, L.norm2, arena + L.conv1_b, NOGAP, NOGAP, P.mid, eps, pk, L.p_o_scale, L.p_o_shift,
                    L.p_o_alpha, st));
      TCVN_TRY(repack(arena + L.conv2_w, d->growth, P.mid, 9, NOGAP, NOGAP, P.mid, d->growth, bf, bf, pk + L.p_w2, st));
      TCVN_TRY(pad_copy(arena + L.conv2_b, d->growth, reinterpret_cast<float*>(pk + L.p_b2), d->growth, st));
    }
    if (B.has_transition) {
      TCVN_TRY(fold(arena, B.tnorm, nullptr, B.c0, B.c0p, B.tkpad, eps, pk, B.p_t_scale, B.p_t_shift, B.p_t_alpha, st));
      TCVN_TRY(repack(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.ctot, B.toutp, false, false, pk + B.p_tw, st));
      TCVN_TRY(pad_copy(arena + B.tconv_b, B.tout, reinterpret_cast<float*>(pk + B.p_tb), B.toutp, st));
      if (bf) {
        const int n16 = B.tn_tiles * 128;
        TCVN_TRY(repack(arena + B.tconv_w, B.tout, B.clog, 1, B.c0, B.c0p, B.tkpad, n16, true, true, pk + B.p_tw16, st));
        TCVN_TRY(pad_copy(arena + B.tconv_b, B.tout, reinterpret_cast<float*>(pk + B.p_tb16), n16, st));
        fill_kernel<<<ceil_div(n16, 128), 128, 0, st>>>(reinterpret_cast<float*>(pk + B.p_ta16), n16, 1.0f);
        TCVN_LAUNCH_CHECK();
      }
    }
  }
  const BlockPlan& last = P.blocks.back();
  TCVN_TRY(fold(arena, P.final_norm, nullptr, last.c0, last.c0p, last.ctot, eps, pk, P.p_f_scale, P.p_f_shift,
                P.p_f_alpha, st));
  TCVN_TRY(repack(arena + P.lin_w, d->out_features, last.clog, 1, last.c0, last.c0p, last.ctot, d->out_features, false,
                  false, pk + P.p_lw, st));
  TCVN_TRY(fold(arena, P.out_norm, nullptr, NOGAP, NOGAP, d->out_features, eps, pk, P.p_lo_scale, P.p_lo_shift,
                P.p_lo_alpha, st));
  return TCVN_OK;
}
