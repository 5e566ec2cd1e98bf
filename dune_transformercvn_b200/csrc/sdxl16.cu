// bf16 / tcgen05 path of the --sdxl pixel-map CNN (BASELINE configs[3]; transformercvn/network/layers/sdxl_net.py:7-42 =
// diffusers' VAE Encoder, parity unpinned - see sdxl.cu / oracle/restate_sdxl.py).  Feature maps are ringed
// channels-last bf16 matrices [n * (H+2) * (W+2)][C]; every 3x3 / 1x1 convolution is the shifted GEMM of umma.cu
// (launch_gemm_shifted: 9 row-shifted views of the activated map as the K loop, the residual / 1x1 shortcut as a trailing
// K segment of the SAME GEMM, bias in the epilogue), so a ResNet block is two tensor-core launches.  This file holds what
// surrounds them:
//   patch27     conv_in (3 -> 64 channels): 3x3 patches of the NCHW fp32 pixels as ONE 64-wide K chunk (27 real columns)
//   groupnorm   GroupNorm(1 group) + SiLU: per-image statistics (per-CTA partial sums in fixed slots, added in a fixed
//               order: bit-reproducible) and the elementwise pass, 8 channels per thread
//   patch_s2    patches of the stride-2 down-sampling convolution behind F.pad(x, (0, 1, 0, 1))
//   to_f32      bf16 ringed map -> fp32 (the 1x1-spatial tail of the network runs the fp32 kernels)
#include "kernels.h"
#include "umma.h"

namespace tcvn {

namespace {

typedef __nv_bfloat16 bf;

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// out[row][(dy*3 + dx)*C + c] = px[img][c][y + dy - 1][x + dx - 1] / divisor (0 outside the map), columns >= 9*C zero, ring rows zero
__global__ void patch27_kernel(const float* __restrict__ px, int n, int C, int H, int W, float divisor, bf* __restrict__ out) {
  const int Hp = H + 2, Wp = W + 2;
  const long long rows = (long long)n * Hp * Wp;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    long long q = r;
    const int x = (int)(q % Wp) - 1; q /= Wp;
    const int y = (int)(q % Hp) - 1;
    const int img = (int)(q / Hp);
    const bool interior = y >= 0 && y < H && x >= 0 && x < W;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    if (interior) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c >= C) break;
        const float* plane = px + ((long long)img * C + c) * H * W;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = y + dy - 1;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int xx = x + dx - 1;
            float t = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) t = __ldg(plane + (long long)yy * W + xx);
            if (divisor != 0.f) t = __fdiv_rn(t, divisor);
            v[(dy * 3 + dx) * 3 + c] = t;   // C <= 3 (checked by the launcher): column (tap, channel)
          }
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + r * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float f8[8] = {v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]};
      dst[i] = pack8(f8);
    }
  }
}

// per-image (sum, sum^2) over the whole ringed slab (ring rows are zero): grid (slabs, images), slot [img][slab][2]
constexpr int kGnSlabs = 128;
__global__ void __launch_bounds__(256) gn_stats16_kernel(const bf* __restrict__ x, long long vecs_per_image, double* __restrict__ parts) {
  const int img = blockIdx.y;
  const uint4* base = reinterpret_cast<const uint4*>(x) + (long long)img * vecs_per_image;
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vecs_per_image; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(__ldg(base + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1 += f[k]; s2 = fmaf(f[k], f[k], s2); }
    if (++cnt == 32) { d1 += (double)s1; d2 += (double)s2; s1 = 0.f; s2 = 0.f; cnt = 0; }
  }
  d1 += (double)s1; d2 += (double)s2;
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = d1;
  sh[1][threadIdx.x] = d2;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    parts[((long long)img * gridDim.x + blockIdx.x) * 2] = sh[0][0];
    parts[((long long)img * gridDim.x + blockIdx.x) * 2 + 1] = sh[1][0];
  }
}

// stat[img] = (mean, rstd) from the slab partials, added in slab order
__global__ void gn_finalize16_kernel(const double* __restrict__ parts, int n, int slabs, double count, float eps, float2* __restrict__ stat) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < slabs; ++s) { a += parts[((long long)img * slabs + s) * 2]; b += parts[((long long)img * slabs + s) * 2 + 1]; }
  const double mean = a / count;
  double var = b / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stat[img] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// out = act((x - mean) * rstd * gamma + beta) on interior rows, 0 on ring rows; one thread = 8 channels of a row
__global__ void __launch_bounds__(256) gn_apply16_kernel(const bf* __restrict__ x, int C, int Hp, int Wp, long long rows,
                                                         const float2* __restrict__ stat, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int silu, bf* __restrict__ out) {
  const int cv = C >> 3;
  const long long total = rows * cv;
  const long long rpi = (long long)Hp * Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    const long long r = i / cv;
    const int img = (int)(r / rpi);
    const int rr = (int)(r - (long long)img * rpi);
    const int y = rr / Wp, xx = rr - y * Wp;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
    if (!(y == 0 || y == Hp - 1 || xx == 0 || xx == Wp - 1)) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * C + c)), f);
      const float2 st = __ldg(stat + img);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      // silu(v) = h + h * tanh(h), h = v / 2 (one MUFU per value, see umma_conv2d.cu); packed fp32 pairs
      const float half = silu ? 0.5f : 1.0f;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float2 sc = make_float2(half * st.y * ga[k], half * st.y * ga[k + 1]);
        const float2 sh = make_float2(fmaf(-st.x, sc.x, half * be[k]), fmaf(-st.x, sc.y, half * be[k + 1]));
        float2 h = __ffma2_rn(make_float2(f[k], f[k + 1]), sc, sh);
        if (silu) {
          float tx, ty;
          asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
          h = __ffma2_rn(h, make_float2(tx, ty), h);
        }
        f[k] = h.x; f[k + 1] = h.y;
      }
    }
    *reinterpret_cast<uint4*>(out + r * C + c) = pack8(f);
  }
}

// patches of the stride-2 3x3 convolution behind F.pad(x, (0, 1, 0, 1)), output geometry ringed, K = 9*C (tap-major)
__global__ void __launch_bounds__(256) patch_s2_16_kernel(const bf* __restrict__ x, int n, int C, int H, int W, int Ho, int Wo,
                                                          bf* __restrict__ out) {
  const int Wp = W + 2, Hp = H + 2, Wop = Wo + 2, Hop = Ho + 2;
  const int cv = C >> 3;
  const long long total = (long long)n * Hop * Wop * 9 * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int tap = (int)(r % 9); r /= 9;
    const long long orow = r;
    const int xo = (int)(r % Wop) - 1; r /= Wop;
    const int yo = (int)(r % Hop) - 1;
    const int img = (int)(r / Hop);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yo >= 0 && yo < Ho && xo >= 0 && xo < Wo) {
      const int y = 2 * yo + tap / 3, xx = 2 * xo + tap % 3;
      if (y < H && xx < W) v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)img * Hp + y + 1) * Wp + xx + 1) * C + c));
    }
    *reinterpret_cast<uint4*>(out + (orow * 9 + tap) * C + c) = v;
  }
}

// space-to-depth input of the stride-2 3x3 convolution behind F.pad(x, (0, 1, 0, 1)): out has the OUTPUT's ringed geometry
// [n][(Ho+2)(Wo+2)][4*C]; column group g = 2*py + px of ringed position (i+1, j+1) holds x[2i+py][2j+px] (0 outside the map;
// the bottom / right ring positions carry the row 2*Ho / column 2*Wo an odd-sized map still has).  Tap (dy, dx) of the
// convolution is then group (dy & 1, dx & 1) shifted by (dy >> 1, dx >> 1): nine row-shifted views of ONE matrix, 512 B
// per output pixel instead of the 1152 B of materialised patches.
__global__ void __launch_bounds__(256) s2d_16_kernel(const bf* __restrict__ x, int n, int C, int H, int W, int Ho, int Wo,
                                                     bf* __restrict__ out) {
  const int Wp = W + 2, Hp = H + 2, Wop = Wo + 2, Hop = Ho + 2;
  const int cv = C >> 3;
  const long long total = (long long)n * Hop * Wop * 4 * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * 8;
    long long r = idx / cv;
    const int g = (int)(r & 3); r >>= 2;
    const long long orow = r;
    const int j = (int)(r % Wop) - 1; r /= Wop;
    const int i = (int)(r % Hop) - 1;
    const int img = (int)(r / Hop);
    const int y = 2 * i + (g >> 1), xx = 2 * j + (g & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i >= 0 && j >= 0 && y < H && xx < W) v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)img * Hp + y + 1) * Wp + xx + 1) * C + c));
    *reinterpret_cast<uint4*>(out + (orow * 4 + g) * C + c) = v;
  }
}

__global__ void bf16_to_f32_kernel(const bf* __restrict__ x, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(x[i]);
}

inline int grid_for(long long total) {
  long long b = ceil_div_ll(total, 256);
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace tcvn

using namespace tcvn;

extern "C" int tcvn_sdxl16_patch27(const float* pixels_nchw, int n, int C, int H, int W, float divisor, void* out_bf16,
                                   tcvn_stream_t stream) {
  TCVN_CHECK_ARG(pixels_nchw && out_bf16 && n >= 0 && C >= 1 && C <= 3 && H > 0 && W > 0, "sdxl16_patch27: bad arguments (C %d)", C);
  if (n == 0) return TCVN_OK;
  patch27_kernel<<<grid_for((long long)n * (H + 2) * (W + 2)), 256, 0, stream>>>(pixels_nchw, n, C, H, W, divisor, static_cast<bf*>(out_bf16));
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" size_t tcvn_sdxl16_groupnorm_workspace_bytes(int n) { return (size_t)n * (kGnSlabs * 2 * sizeof(double) + sizeof(float2)) + 256; }

extern "C" int tcvn_sdxl16_groupnorm(const void* x_bf16, int n, int C, int H, int W, const float* gamma, const float* beta, float eps,
                                     int silu, void* out_bf16, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && gamma && beta && out_bf16 && workspace && n >= 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0,
                 "sdxl16_groupnorm: bad arguments (C %d)", C);
  if (n == 0) return TCVN_OK;
  if (workspace_bytes < tcvn_sdxl16_groupnorm_workspace_bytes(n))
    return fail(TCVN_ERR_WORKSPACE, "sdxl16_groupnorm: workspace %zu < %zu bytes", workspace_bytes, tcvn_sdxl16_groupnorm_workspace_bytes(n));
  const long long rpi = (long long)(H + 2) * (W + 2);
  const long long vecs = rpi * C / 8;
  int slabs = (int)ceil_div_ll(vecs, 256 * 8);
  if (slabs > kGnSlabs) slabs = kGnSlabs;
  double* parts = static_cast<double*>(workspace);
  float2* stat = reinterpret_cast<float2*>(parts + (size_t)n * kGnSlabs * 2);
  gn_stats16_kernel<<<dim3(slabs, n), 256, 0, stream>>>(static_cast<const bf*>(x_bf16), vecs, parts);
  TCVN_LAUNCH_CHECK();
  gn_finalize16_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(parts, n, slabs, (double)H * W * C, eps, stat);
  TCVN_LAUNCH_CHECK();
  gn_apply16_kernel<<<grid_for((long long)n * rpi * (C / 8)), 256, 0, stream>>>(static_cast<const bf*>(x_bf16), C, H + 2, W + 2,
                                                                               (long long)n * rpi, stat, gamma, beta, silu,
                                                                               static_cast<bf*>(out_bf16));
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// stat[img] = (mean, rstd) of GroupNorm(1 group) over a ringed bf16 map, without the elementwise pass (the consumer applies
// it on operand load: tcvn_sdxl16_conv2d_c64).  workspace: tcvn_sdxl16_groupnorm_workspace_bytes(n)
extern "C" int tcvn_sdxl16_gn_stats(const void* x_bf16, int n, int C, int H, int W, float eps, void* out_stat, void* workspace,
                                    size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && out_stat && workspace && n >= 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, "sdxl16_gn_stats: bad arguments");
  if (n == 0) return TCVN_OK;
  if (workspace_bytes < tcvn_sdxl16_groupnorm_workspace_bytes(n))
    return fail(TCVN_ERR_WORKSPACE, "sdxl16_gn_stats: workspace %zu < %zu bytes", workspace_bytes, tcvn_sdxl16_groupnorm_workspace_bytes(n));
  const long long vecs = (long long)(H + 2) * (W + 2) * C / 8;
  int slabs = (int)ceil_div_ll(vecs, 256 * 8);
  if (slabs > kGnSlabs) slabs = kGnSlabs;
  double* parts = static_cast<double*>(workspace);
  gn_stats16_kernel<<<dim3(slabs, n), 256, 0, stream>>>(static_cast<const bf*>(x_bf16), vecs, parts);
  TCVN_LAUNCH_CHECK();
  gn_finalize16_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(parts, n, slabs, (double)H * W * C, eps, static_cast<float2*>(out_stat));
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_sdxl16_patch_s2(const void* x_bf16, int n, int C, int H, int W, void* out_bf16, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && out_bf16 && n >= 0 && C > 0 && C % 8 == 0 && H >= 2 && W >= 2, "sdxl16_patch_s2: bad arguments");
  if (n == 0) return TCVN_OK;
  const int Ho = H / 2, Wo = W / 2;
  patch_s2_16_kernel<<<grid_for((long long)n * (Ho + 2) * (Wo + 2) * 9 * (C / 8)), 256, 0, stream>>>(
      static_cast<const bf*>(x_bf16), n, C, H, W, Ho, Wo, static_cast<bf*>(out_bf16));
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_sdxl16_s2d(const void* x_bf16, int n, int C, int H, int W, void* out_bf16, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && out_bf16 && n >= 0 && C > 0 && C % 64 == 0 && H >= 2 && W >= 2, "sdxl16_s2d: bad arguments");
  if (n == 0) return TCVN_OK;
  const int Ho = H / 2, Wo = W / 2;
  s2d_16_kernel<<<grid_for((long long)n * (Ho + 2) * (Wo + 2) * 4 * (C / 8)), 256, 0, stream>>>(
      static_cast<const bf*>(x_bf16), n, C, H, W, Ho, Wo, static_cast<bf*>(out_bf16));
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// stride-2 3x3 convolution over the space-to-depth matrix of tcvn_sdxl16_s2d: out[rows][out_cols] in the output's ringed
// geometry (ring_hp x ring_wp), ring rows zero; w_bf16 [n_tiles * 128][9 * C] tap-major like tcvn_sdxl16_conv
extern "C" int tcvn_sdxl16_conv_s2(const void* s2d_bf16, int64_t rows, int C, const void* w_bf16, int n_tiles, const float* bias_padded,
                                   const float* ones_padded, void* out_bf16, int out_cols, int ring_hp, int ring_wp,
                                   tcvn_stream_t stream) {
  TCVN_CHECK_ARG(s2d_bf16 && w_bf16 && bias_padded && ones_padded && out_bf16 && rows >= 0 && C > 0 && C % 64 == 0 && n_tiles >= 1 &&
                     ring_hp >= 3 && ring_wp >= 3 && out_cols % 8 == 0, "sdxl16_conv_s2: bad arguments");
  if (rows == 0) return TCVN_OK;
  int tap_off[9], tap_col[9];
  for (int t = 0; t < 9; ++t) {
    const int dy = t / 3, dx = t % 3;
    tap_off[t] = (dy >> 1) * ring_wp + (dx >> 1);
    tap_col[t] = ((dy & 1) * 2 + (dx & 1)) * C;
  }
  return launch_gemm_shifted(s2d_bf16, rows, C, 9, tap_off, nullptr, 0, w_bf16, n_tiles, bias_padded, ones_padded, out_bf16, out_cols,
                             ring_hp, ring_wp, stream, 4 * C, tap_col);
}

extern "C" int tcvn_sdxl16_to_f32(const void* x_bf16, int64_t count, float* out, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && out && count >= 0, "sdxl16_to_f32: bad arguments");
  if (count == 0) return TCVN_OK;
  bf16_to_f32_kernel<<<grid_for(count), 256, 0, stream>>>(static_cast<const bf*>(x_bf16), count, out);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// out[m, 0:out_cols] = sum_t A[m + tap_off[t], :] W_t^T + X2[m, :] W_x^T + bias ; ring rows of out zero (umma.cu)
extern "C" int tcvn_sdxl16_conv(const void* a_bf16, int64_t rows, int a_cols, int n_taps, const int32_t* tap_off, const void* x2_bf16,
                                int x2_cols, const void* w_bf16, int n_tiles, const float* bias_padded, const float* ones_padded,
                                void* out_bf16, int out_cols, int ring_hp, int ring_wp, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(a_bf16 && w_bf16 && bias_padded && ones_padded && out_bf16 && rows >= 0 && n_tiles >= 1 && ring_hp >= 3 && ring_wp >= 3 &&
                     out_cols % 8 == 0 && (n_taps == 1 || tap_off),
                 "sdxl16_conv: bad arguments");
  if (rows == 0) return TCVN_OK;
  return launch_gemm_shifted(a_bf16, rows, a_cols, n_taps, tap_off, x2_bf16, x2_cols, w_bf16, n_tiles, bias_padded, ones_padded, out_bf16,
                             out_cols, ring_hp, ring_wp, stream);
}
