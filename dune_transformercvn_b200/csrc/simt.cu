// CUDA-core kernels of the hot path: the fp32 parity path's GEMM / shifted-GEMM convolution, the
// stem, the pooling / BN+PReLU epilogue kernels shared by both precisions.
//
// Reference arithmetic (transformercvn/network/layers/dense_net.py):
//   stem        :111-122  conv7x7 s2 p3 + bias -> BN -> PReLU -> AvgPool2d(3, 2)
//   Bottleneck  :8-45     BN -> PReLU -> conv1x1 -> BN -> PReLU -> conv3x3 p1 -> cat
//   Transition  :78-94    BN -> PReLU -> conv1x1 -> AvgPool2d(2, 2)
//   tail        :147-162  BN -> PReLU -> AdaptiveAvgPool2d(1) -> Linear -> BN1d -> PReLU
// Eval mode: every BatchNorm is the affine map folded by pack.cu.
#include "kernels.h"

namespace tcvn {

// ------------------------------------------------------------------------------------------------
// Shifted GEMM on CUDA cores (fp32 FMA).  64 x BN output tile, 256 threads, 4 x TN per thread.
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBK = 16;

struct GemmDev {
  const void* A; int lda; long long m_total; int K; int taps; int tap_off[9];
  const float* W; int N;
  const float* a_scale; const float* a_shift; const float* a_alpha;
  const float* o_scale; const float* o_shift; const float* o_alpha;
  void* out; int ldo; int out_col0;
  int ring_Hp, ring_Wp;
};

template <typename TA, typename TO, int TN>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const GemmDev g) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][BN];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const TA* A = static_cast<const TA*>(g.A);
  const int ar = tid >> 2, akq = (tid & 3) * 4;          // A loader: row, first k of a group of 4
  constexpr int BT = BN / 4;                             // B loader threads per k-row
  const int bk = tid / BT, bn = (tid % BT) * 4;
  const bool transform = g.a_scale != nullptr;

  for (int t = 0; t < g.taps; ++t) {
    const long long gm = m0 + ar + g.tap_off[t];
    const bool row_ok = gm >= 0 && gm < g.m_total;
    const TA* arow = A + (row_ok ? gm : 0) * (long long)g.lda;
    const float* Wt = g.W + (size_t)t * g.K * g.N;
    for (int k0 = 0; k0 < g.K; k0 += kBK) {
      float av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + akq + i;
        float v = 0.f;
        if (row_ok && k < g.K) {
          v = to_f32<TA>(arow[k]);
          if (transform) v = prelu(fmaf(v, __ldg(g.a_scale + k), __ldg(g.a_shift + k)), __ldg(g.a_alpha + k));
        }
        av[i] = v;
      }
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bk < kBK && k0 + bk < g.K && n0 + bn < g.N)
        bv = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + bk) * g.N + n0 + bn));
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) As[akq + i][ar] = av[i];
      if (bk < kBK) *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  TO* out = static_cast<TO*>(g.out);
  const int R = g.ring_Hp * g.ring_Wp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= g.m_total) continue;
    bool ring = false;
    if (R > 0) {
      const int rr = (int)(m % R);
      const int y = rr / g.ring_Wp, x = rr - y * g.ring_Wp;
      ring = y == 0 || y == g.ring_Hp - 1 || x == 0 || x == g.ring_Wp - 1;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.o_scale != nullptr) v = prelu(fmaf(v, __ldg(g.o_scale + n), __ldg(g.o_shift + n)), __ldg(g.o_alpha + n));
      else v += __ldg(g.o_shift + n);
      if (ring) v = 0.f;
      out[m * (long long)g.ldo + g.out_col0 + n] = from_f32<TO>(v);
    }
  }
}

int launch_simt_gemm(const GemmArgs& a, cudaStream_t stream) {
  TCVN_CHECK_ARG(a.N % 4 == 0 && a.taps >= 1 && a.taps <= 9 && a.o_shift != nullptr, "simt_gemm: bad arguments");
  if (a.m_total <= 0) return TCVN_OK;
  GemmDev g;
  g.A = a.A; g.lda = a.lda; g.m_total = a.m_total; g.K = a.K; g.taps = a.taps;
  for (int i = 0; i < 9; ++i) g.tap_off[i] = a.tap_off[i];
  g.W = a.W; g.N = a.N;
  g.a_scale = a.a_scale; g.a_shift = a.a_shift; g.a_alpha = a.a_alpha;
  g.o_scale = a.o_scale; g.o_shift = a.o_shift; g.o_alpha = a.o_alpha;
  g.out = a.out; g.ldo = a.ldo; g.out_col0 = a.out_col0; g.ring_Hp = a.ring_Hp; g.ring_Wp = a.ring_Wp;
  const bool narrow = a.N <= 32;
  const int BN = narrow ? 32 : 64;
  dim3 grid((unsigned)ceil_div_ll(a.m_total, kBM), (unsigned)ceil_div(a.N, BN));
#define TCVN_GEMM_CASE(TA, TO)                                                     \
  do {                                                                             \
    if (narrow) simt_gemm_kernel<TA, TO, 2><<<grid, 256, 0, stream>>>(g);          \
    else simt_gemm_kernel<TA, TO, 4><<<grid, 256, 0, stream>>>(g);                 \
  } while (0)
  if (a.a_is_f32 && a.out_is_f32) TCVN_GEMM_CASE(float, float);
  else if (!a.a_is_f32 && !a.out_is_f32) TCVN_GEMM_CASE(__nv_bfloat16, __nv_bfloat16);
  else if (a.a_is_f32 && !a.out_is_f32) TCVN_GEMM_CASE(float, __nv_bfloat16);
  else TCVN_GEMM_CASE(__nv_bfloat16, float);
#undef TCVN_GEMM_CASE
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// ------------------------------------------------------------------------------------------------
// Stem, fused: 7x7 stride-2 convolution on NCHW fp32 pixels + bias + BN0 + PReLU0 + AvgPool2d(3, 2),
// written straight into channels [0, 64) of block 0's ringed buffer (dense_net.py:111-122).  The
// 64 x 200 x 140 pre-pool map (3.6 MB/image in bf16 - the largest tensor of the network) never
// reaches memory.  Persistent CTAs keep the 147 x 64 filter bank in shared memory and loop over
// 8 x 8 tiles of pooled pixels (17 x 17 conv outputs, 39 x 39 x 3 input window).  One thread per
// conv output, 64 channels in registers; the pixel maps are ~99 % zeros, so a warp skips a tap when
// none of its 32 input values is non-zero (adding 0*w is exact: skipping is bit-identical).
// ------------------------------------------------------------------------------------------------
constexpr int kStemTP = 8;                   // pooled tile edge
constexpr int kStemTC = 2 * kStemTP + 1;     // 17 conv outputs per edge
constexpr int kStemIn = 2 * kStemTC + 5;     // 39 input pixels per edge
constexpr int kStemThreads = 320;            // 289 conv outputs -> 10 warps

template <typename TO, int C0>
__global__ void __launch_bounds__(kStemThreads) stem_fused_kernel(const float* __restrict__ pixels, int n_images, int cin,
                                                                  int H, int W, int Hs, int Ws,
                                                                  const float* __restrict__ w0,
                                                                  const float* __restrict__ s_scale,
                                                                  const float* __restrict__ s_shift,
                                                                  const float* __restrict__ s_alpha,
                                                                  TO* __restrict__ blk, int ldo, int Hb, int Wb) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;                                        // [cin*49][C0]
  float* ism = wsm + cin * 49 * C0;                         // [cin][39][40]
  TO* csm = reinterpret_cast<TO*>(ism + cin * kStemIn * (kStemIn + 1));  // [289][C0] activated conv outputs
  for (int i = threadIdx.x; i < cin * 49 * C0; i += blockDim.x) wsm[i] = __ldg(w0 + i);
  const int tiles_x = (Wb + kStemTP - 1) / kStemTP, tiles_y = (Hb + kStemTP - 1) / kStemTP;
  const int per_image = tiles_x * tiles_y;
  const long long total = (long long)n_images * per_image;
  const int t = threadIdx.x;
  const bool active = t < kStemTC * kStemTC;
  const int cy = active ? t / kStemTC : 0, cx = active ? t % kStemTC : 0;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = (int)(tile / per_image);
    const int rem = (int)(tile - (long long)n * per_image);
    const int py0 = (rem / tiles_x) * kStemTP, px0 = (rem % tiles_x) * kStemTP;
    const int iy0 = 4 * py0 - 3, ix0 = 4 * px0 - 3;
    const float* img = pixels + (size_t)n * cin * H * W;
    __syncthreads();  // previous tile's pooling reads of csm / conv reads of ism are done
    for (int i = t; i < cin * kStemIn * kStemIn; i += blockDim.x) {
      const int c = i / (kStemIn * kStemIn);
      const int r = i - c * kStemIn * kStemIn;
      const int yy = r / kStemIn, xx = r - yy * kStemIn;
      const int y = iy0 + yy, x = ix0 + xx;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(img + ((size_t)c * H + y) * W + x);
      ism[(c * kStemIn + yy) * (kStemIn + 1) + xx] = v;
    }
    __syncthreads();
    float acc[C0];
#pragma unroll
    for (int j = 0; j < C0; ++j) acc[j] = 0.f;
    for (int c = 0; c < cin; ++c)
      for (int ky = 0; ky < 7; ++ky) {
        const float* irow = ism + (c * kStemIn + 2 * cy + ky) * (kStemIn + 1) + 2 * cx;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float v = active ? irow[kx] : 0.f;
          if (__any_sync(0xffffffffu, v != 0.f)) {
            const float4* w4 = reinterpret_cast<const float4*>(wsm + ((c * 7 + ky) * 7 + kx) * C0);
#pragma unroll
            for (int j = 0; j < C0 / 4; ++j) {
              const float4 w = w4[j];
              acc[4 * j + 0] = fmaf(v, w.x, acc[4 * j + 0]);
              acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
              acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
              acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
            }
          }
        }
      }
    if (active) {
      TO* o = csm + (size_t)t * C0;
#pragma unroll
      for (int j = 0; j < C0; ++j)
        o[j] = from_f32<TO>(prelu(fmaf(acc[j], __ldg(s_scale + j), __ldg(s_shift + j)), __ldg(s_alpha + j)));
    }
    __syncthreads();
    for (int i = t; i < kStemTP * kStemTP * C0; i += blockDim.x) {
      const int ch = i % C0;
      const int p = i / C0;
      const int pyl = p / kStemTP, pxl = p % kStemTP;
      const int py = py0 + pyl, px = px0 + pxl;
      if (py >= Hb || px >= Wb) continue;
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) s += to_f32<TO>(csm[((2 * pyl + dy) * kStemTC + 2 * pxl + dx) * C0 + ch]);
      const size_t row = (size_t)n * (Hb + 2) * (Wb + 2) + (size_t)(py + 1) * (Wb + 2) + (px + 1);
      blk[row * ldo + ch] = from_f32<TO>(s / 9.0f);
    }
  }
}

int launch_stem(const float* pixels, int n, int cin, int H, int W, const float* w0, const float* s_scale,
                const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb, int Wb, bool f32,
                cudaStream_t stream) {
  if (c0 != 64) return fail(TCVN_ERR_UNSUPPORTED, "stem: init_features %d (kernel is specialised for 64)", c0);
  if (n == 0) return TCVN_OK;
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  if ((Hs - 3) / 2 + 1 != Hb || (Ws - 3) / 2 + 1 != Wb) return fail(TCVN_ERR_ARG, "stem: geometry mismatch");
  const size_t smem = ((size_t)cin * 49 * c0 + (size_t)cin * kStemIn * (kStemIn + 1)) * sizeof(float) +
                      (size_t)kStemTC * kStemTC * c0 * (f32 ? 4 : 2);
  if (smem > 220 * 1024) return fail(TCVN_ERR_UNSUPPORTED, "stem: %d input channels do not fit shared memory", cin);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long tiles = (long long)n * ((Wb + kStemTP - 1) / kStemTP) * ((Hb + kStemTP - 1) / kStemTP);
  const int per_sm = f32 ? 1 : 2;
  const int grid = (int)(tiles < (long long)sms * per_sm ? tiles : (long long)sms * per_sm);
  if (f32) {
    TCVN_CUDA(cudaFuncSetAttribute(stem_fused_kernel<float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_fused_kernel<float, 64><<<grid, kStemThreads, smem, stream>>>(pixels, n, cin, H, W, Hs, Ws, w0, s_scale, s_shift,
                                                                        s_alpha, static_cast<float*>(blk), ldo, Hb, Wb);
  } else {
    TCVN_CUDA(cudaFuncSetAttribute(stem_fused_kernel<__nv_bfloat16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    stem_fused_kernel<__nv_bfloat16, 64><<<grid, kStemThreads, smem, stream>>>(
        pixels, n, cin, H, W, Hs, Ws, w0, s_scale, s_shift, s_alpha, static_cast<__nv_bfloat16*>(blk), ldo, Hb, Wb);
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// Transition front half.  AvgPool2d(2,2) is linear and the 1x1 convolution is per-pixel, so
// pool(conv(a)) == conv(pool(a)) (bias included): pooling the activated map first makes the GEMM 4x smaller.
template <typename T>
__global__ void act_pool2_kernel(const T* __restrict__ blk, int H, int W, int ld, int c,
                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ alpha, T* __restrict__ out, int H2, int W2,
                                 long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  long long r = idx / c;
  const int x = (int)(r % W2); r /= W2;
  const int y = (int)(r % H2);
  const int n = (int)(r / H2);
  const int Wp = W + 2;
  const size_t row0 = (size_t)n * (H + 2) * Wp + (size_t)(2 * y + 1) * Wp + (2 * x + 1);
  const float sc = __ldg(scale + ch), sh = __ldg(shift + ch), al = __ldg(alpha + ch);
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
      s += prelu(fmaf(to_f32<T>(blk[(row0 + (size_t)dy * Wp + dx) * ld + ch]), sc, sh), al);
  const size_t orow = (size_t)n * (H2 + 2) * (W2 + 2) + (size_t)(y + 1) * (W2 + 2) + (x + 1);
  out[orow * c + ch] = from_f32<T>(s * 0.25f);
}

int launch_act_pool2(const void* blk, int n, int H, int W, int ld, int c, const float* scale, const float* shift,
                     const float* alpha, void* out, int H2, int W2, bool f32, cudaStream_t stream) {
  const long long total = (long long)n * H2 * W2 * c;
  if (total == 0) return TCVN_OK;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  if (f32) act_pool2_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(blk), H, W, ld, c, scale, shift,
                                                             alpha, static_cast<float*>(out), H2, W2, total);
  else act_pool2_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(blk), H, W, ld, c,
                                                                 scale, shift, alpha, static_cast<__nv_bfloat16*>(out),
                                                                 H2, W2, total);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

template <typename T>
__global__ void act_gap_kernel(const T* __restrict__ blk, int H, int W, int ld, int c, const float* __restrict__ scale,
                               const float* __restrict__ shift, const float* __restrict__ alpha,
                               float* __restrict__ gap) {
  const int n = blockIdx.x;
  const int Wp = W + 2;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float sc = __ldg(scale + ch), sh = __ldg(shift + ch), al = __ldg(alpha + ch);
    float s = 0.f;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t row = (size_t)n * (H + 2) * Wp + (size_t)(y + 1) * Wp + (x + 1);
        s += prelu(fmaf(to_f32<T>(blk[row * ld + ch]), sc, sh), al);
      }
    gap[(size_t)n * c + ch] = s / (float)(H * W);
  }
}

int launch_act_gap(const void* blk, int n, int H, int W, int ld, int c, const float* scale, const float* shift,
                   const float* alpha, float* gap, bool f32, cudaStream_t stream) {
  if (n == 0) return TCVN_OK;
  if (f32) act_gap_kernel<float><<<n, 128, 0, stream>>>(static_cast<const float*>(blk), H, W, ld, c, scale, shift, alpha, gap);
  else act_gap_kernel<__nv_bfloat16><<<n, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(blk), H, W, ld, c, scale,
                                                            shift, alpha, gap);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// test hook: ringed channels-last -> NCHW fp32, dropping the `c_skip` alignment-padding channels at c_skip_from
template <typename T>
__global__ void ring_to_nchw_kernel(const T* __restrict__ blk, int H, int W, int ld, int c, int c_skip_from, int c_skip,
                                    float* __restrict__ out, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long r = idx / W;
  const int y = (int)(r % H); r /= H;
  const int ch = (int)(r % c);
  const int n = (int)(r / c);
  const int pc = ch < c_skip_from ? ch : ch + c_skip;
  const size_t row = (size_t)n * (H + 2) * (W + 2) + (size_t)(y + 1) * (W + 2) + (x + 1);
  out[idx] = to_f32<T>(blk[row * ld + pc]);
}

int launch_ring_to_nchw(const void* blk, int n, int H, int W, int ld, int c, int c_skip_from, int c_skip, float* out,
                        bool f32, cudaStream_t stream) {
  const long long total = (long long)n * c * H * W;
  if (total == 0) return TCVN_OK;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  if (f32) ring_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(blk), H, W, ld, c, c_skip_from,
                                                                c_skip, out, total);
  else ring_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(blk), H, W, ld, c,
                                                                    c_skip_from, c_skip, out, total);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn
