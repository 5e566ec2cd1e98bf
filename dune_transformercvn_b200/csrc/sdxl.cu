// Kernels of the --sdxl pixel-map CNN (BASELINE configs[3]) that the DenseNet path does not have:
// GroupNorm(1 group) + SiLU, the stride-2 patch gather of the down-sampling convolution, NCHW -> ringed channels-last.
// The convolutions themselves are the shifted GEMM of the fp32 parity path (tcvn_t_gemm).
//
// Reference: transformercvn/network/layers/sdxl_net.py:7-42 builds diffusers.models.vae.Encoder (un-vendored third-party
// code, no version pinned; SURVEY 8c: parity unpinned).  Arithmetic restated from diffusers' published layout:
//   ResnetBlock2D   h = conv1(silu(GroupNorm(x))); h = conv2(silu(GroupNorm(h))); out = shortcut(x) + h      (eps 1e-6)
//   Downsample2D    F.pad(x, (0, 1, 0, 1)) -> conv3x3 stride 2 padding 0
//   GroupNorm       per sample and group: (x - mean) / sqrt(biased var + eps) * gamma_c + beta_c
#include "common.cuh"

namespace tcvn {

// NCHW fp32 pixels -> interior of the ringed channels-last matrix [n][(H+2)(W+2)][C]; ring rows are written as zero
__global__ void pixels_to_ring_kernel(const float* __restrict__ px, int n, int C, int H, int W, float* __restrict__ out) {
  const int Hp = H + 2, Wp = W + 2;
  const long long total = (long long)n * Hp * Wp * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % Wp) - 1; r /= Wp;
    const int y = (int)(r % Hp) - 1;
    const int img = (int)(r / Hp);
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = px[(((long long)img * C + c) * H + y) * W + x];
    out[i] = v;
  }
}

// per (image, group) sum and sum of squares over the interior pixels of a ringed map (ring rows are zero, so the whole
// image slab can be summed); grid = (slabs, groups, images), double atomics into sums[img][group][2]
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int C, int groups, long long rows_per_image,
                                                       double* __restrict__ sums) {
  const int img = blockIdx.z, g = blockIdx.y;
  const int cg = C / groups;
  const float* base = x + (long long)img * rows_per_image * C;
  const long long total = rows_per_image * cg;   // elements of this group in this image
  double s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int c = g * cg + (int)(i - r * cg);
    const float v = base[r * C + c];
    s1 += v;
    s2 += (double)v * v;
  }
  __shared__ double sh1[8], sh2[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh1[w]; b += sh2[w]; }
    atomicAdd(sums + ((long long)img * groups + g) * 2, a);
    atomicAdd(sums + ((long long)img * groups + g) * 2 + 1, b);
  }
}

// out = act((x - mean) * rstd * gamma + beta) on interior rows, 0 on ring rows; act = SiLU or identity
__global__ void gn_apply_kernel(const float* __restrict__ x, int n, int C, int groups, int Hp, int Wp, double count,
                                const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu, float* __restrict__ out) {
  const long long rows_per_image = (long long)Hp * Wp;
  const long long total = (long long)n * rows_per_image * C;
  const int cg = C / groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int img = (int)(r / rows_per_image);
    const int rr = (int)(r - (long long)img * rows_per_image);
    const int y = rr / Wp, xx = rr - y * Wp;
    float v = 0.f;
    if (!(y == 0 || y == Hp - 1 || xx == 0 || xx == Wp - 1)) {
      const double* s = sums + ((long long)img * groups + c / cg) * 2;
      const double mean = s[0] / count;
      double var = s[1] / count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      v = (x[i] - (float)mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (silu) v = v / (1.f + expf(-v));
    }
    out[i] = v;
  }
}

// patches of the stride-2 3x3 convolution behind F.pad(x, (0, 1, 0, 1)):
//   out[img][(yo+1)(Wo+2) + xo+1][(dy*3 + dx)*C + c] = x[img][2*yo + dy][2*xo + dx][c]   (0 beyond the map: the pad)
// in the ringed layout of the OUTPUT geometry, so the convolution becomes a plain GEMM with K = 9*C; ring rows zero
__global__ void patch_s2_kernel(const float* __restrict__ x, int n, int C, int H, int W, int Ho, int Wo,
                                float* __restrict__ out) {
  const int Wp = W + 2, Hp = H + 2, Wop = Wo + 2, Hop = Ho + 2;
  const long long total = (long long)n * Hop * Wop * 9 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int tap = (int)(r % 9); r /= 9;
    const int xo = (int)(r % Wop) - 1; r /= Wop;
    const int yo = (int)(r % Hop) - 1;
    const int img = (int)(r / Hop);
    float v = 0.f;
    if (yo >= 0 && yo < Ho && xo >= 0 && xo < Wo) {
      const int y = 2 * yo + tap / 3, xx = 2 * xo + tap % 3;
      if (y < H && xx < W) v = x[(((long long)img * Hp + y + 1) * Wp + xx + 1) * C + c];
    }
    out[i] = v;
  }
}

static inline int grid_for(long long total) {
  long long b = ceil_div_ll(total, 256);
  if (b > 148 * 32) b = 148 * 32;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace tcvn

using namespace tcvn;

extern "C" int tcvn_sdxl_pixels_to_ring(const float* pixels_nchw, int n, int C, int H, int W, float* out_ring,
                                        tcvn_stream_t stream) {
  TCVN_CHECK_ARG(pixels_nchw && out_ring && n >= 0 && C > 0 && H > 0 && W > 0, "sdxl_pixels_to_ring: bad arguments");
  if (n == 0) return TCVN_OK;
  pixels_to_ring_kernel<<<grid_for((long long)n * (H + 2) * (W + 2) * C), 256, 0, stream>>>(pixels_nchw, n, C, H, W, out_ring);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_sdxl_groupnorm(const float* x_ring, int n, int C, int groups, int H, int W, const float* gamma,
                                   const float* beta, float eps, int silu, float* out_ring, double* sums_workspace,
                                   tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_ring && gamma && beta && out_ring && sums_workspace && n >= 0 && C > 0 && groups > 0 && C % groups == 0 &&
                     H > 0 && W > 0,
                 "sdxl_groupnorm: bad arguments (C %d, groups %d)", C, groups);
  if (n == 0) return TCVN_OK;
  const long long rows = (long long)(H + 2) * (W + 2);
  TCVN_CUDA(cudaMemsetAsync(sums_workspace, 0, sizeof(double) * 2 * n * groups, stream));
  long long slabs = ceil_div_ll(rows * (C / groups), 256 * 16);
  if (slabs > 512) slabs = 512;
  dim3 grid((unsigned)slabs, (unsigned)groups, (unsigned)n);
  gn_stats_kernel<<<grid, 256, 0, stream>>>(x_ring, C, groups, rows, sums_workspace);
  TCVN_LAUNCH_CHECK();
  const double count = (double)H * W * (C / groups);
  gn_apply_kernel<<<grid_for((long long)n * rows * C), 256, 0, stream>>>(x_ring, n, C, groups, H + 2, W + 2, count, sums_workspace,
                                                                         gamma, beta, eps, silu, out_ring);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_sdxl_patch_s2(const float* x_ring, int n, int C, int H, int W, float* out_ring, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_ring && out_ring && n >= 0 && C > 0 && H >= 2 && W >= 2, "sdxl_patch_s2: bad arguments");
  if (n == 0) return TCVN_OK;
  const int Ho = H / 2, Wo = W / 2;   // floor((H + 1 - 3) / 2) + 1
  patch_s2_kernel<<<grid_for((long long)n * (Ho + 2) * (Wo + 2) * 9 * C), 256, 0, stream>>>(x_ring, n, C, H, W, Ho, Wo, out_ring);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
