// Stem of the TRAINING path (dense_net.py:112-118 under autograd): raw conv0 output + its batch statistics, and conv0's
// weight gradient.  Both are hit-driven (the pixel maps are 0.2-1 % occupied) and bit-reproducible: every accumulator has
// exactly one owner, the summation orders are fixed by the geometry, and per-CTA partial results land in fixed slots that
// a later kernel adds in a fixed order.  (Round 1 scattered hits with float atomics: the one-ulp run-to-run differences
// of z0 were amplified by the 67 train-mode BatchNorms to 20 % of the gradient.)
//
// Round 2, second version: like the inference stem (stem_coo.cu) the work unit belongs to ONE WARP - no CTA barriers, no
// accumulator zeroing, lane l owns channels 2l, 2l+1:
//   forward   a warp takes an 8 x 8 tile of conv outputs, reads the tile's 21 x 21 input window from the dense pixel map
//             (63 loads in flight), and every non-zero pixel - found by ballot, row by row, left to right - adds v * w[tap] to
//             the <= 4 x 4 outputs it reaches; a 64-bit mask of touched outputs replaces the zeroing pass.
//             z0[n, cy, cx, :] = bf16(b + conv) for every output of the tile, and the tile's (sum z, sum z^2) - the untouched
//             outputs are count x the constant, exactly - go to the CTA's slot.
//   wgrad     a warp takes one input row of one image, finds its non-zero pixels the same way and adds
//             x[c] * dz[n, cy, cx, :] into its own [cin*49][64] accumulator in shared memory; the warps of a CTA are added
//             in warp order into the CTA's slot.
#include "kernels.h"
#include "stem.cuh"

namespace tcvn {

namespace {

constexpr int kFT = 8;                  // conv outputs per tile edge
constexpr int kFWin = 2 * kFT + 5;      // 21 input pixels per edge
constexpr int kFWarps = 10;             // forward: warps per CTA (16 KB of accumulators each)
constexpr int kGWarps = 6;              // wgrad: warps per CTA (37 KB of accumulators each)
constexpr int kGChunks = 12;            // wgrad: 32-pixel chunks of an input row held in registers (W <= 384)

struct StemTrainArgs {
  const float* pixels; int n_images, cin, H, W, Hs, Ws;
  const float* w0;        // [cin*49][C0]           (forward)
  const float* bias;      // [C0]                   (forward)
  __nv_bfloat16* z0;      // [n, Hs, Ws, C0] bf16   (forward)
  double* stat_parts;     // [grid][2][C0]          (forward)
  const __nv_bfloat16* dz; // [n, Hs, Ws, C0] bf16  (wgrad)
  float* dw_parts;        // [grid][cin*49][C0]     (wgrad)
};

__device__ __forceinline__ __nv_bfloat162 bf2(float2 v) { return __float22bfloat162_rn(v); }

// All non-zero pixels of one window row (ballot m, lane = window column), left to right: each adds v * w[tap] to the <= 4 x 4
// conv outputs it reaches.  NOT inlined: the caller's row loop is unrolled 21 times (the window lives in registers), and 21
// copies of this body made the kernel instruction-cache-bound (ncu: no_instruction was the top stall).
__device__ __noinline__ unsigned long long stem_scatter_row(unsigned m, float r0, float r1, float r2, int yy, unsigned long long mask,
                                                            float2* acc, const float4* wA, const float2* wB) {
  const int cy_lo = yy > 6 ? (yy - 5) >> 1 : 0, cy_hi = min(kFT - 1, yy >> 1);
  while (m) {
    const int xx = __ffs((int)m) - 1;
    m &= m - 1u;
    const float v0 = __shfl_sync(0xffffffffu, r0, xx), v1 = __shfl_sync(0xffffffffu, r1, xx), v2 = __shfl_sync(0xffffffffu, r2, xx);
    const float2 vv0 = make_float2(v0, v0), vv1 = make_float2(v1, v1), vv2 = make_float2(v2, v2);
    const int cx_lo = xx > 6 ? (xx - 5) >> 1 : 0, cx_hi = min(kFT - 1, xx >> 1);
    const int ncx = cx_hi - cx_lo + 1;                 // 1..4
    const unsigned long long run = (1ull << ncx) - 1ull;
    for (int cy = cy_lo; cy <= cy_hi; ++cy) {
      const int pos0 = cy * kFT + cx_lo;
      const int tap0 = (yy - 2 * cy) * 7 + xx - 2 * cx_lo;
      const unsigned field = (unsigned)(mask >> pos0);   // bit j: output j of the run was touched
      float2 s[4];
      float4 wa[4];
      float2 wb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = min(j, ncx - 1);                  // a slot beyond the run repeats a legal address, is not stored
        wa[j] = wA[(tap0 - 2 * jj) * 32];
        wb[j] = wB[(tap0 - 2 * jj) * 32];
        s[j] = make_float2(0.f, 0.f);
        if ((field >> jj) & 1u) s[j] = acc[(pos0 + jj) * 32];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = __ffma2_rn(vv0, make_float2(wa[j].x, wa[j].y), s[j]);
        s[j] = __ffma2_rn(vv1, make_float2(wa[j].z, wa[j].w), s[j]);
        s[j] = __ffma2_rn(vv2, wb[j], s[j]);
        if (j < ncx) acc[(pos0 + j) * 32] = s[j];
      }
      mask |= run << pos0;
    }
  }
  return mask;
}

__global__ void __launch_bounds__(kFWarps * 32, 1) stem_train_fwd_kernel(const StemTrainArgs a) {
  constexpr int C0 = 64;
  extern __shared__ __align__(16) float smem[];
  // filter bank per (tap, lane): {w[c0][2l], w[c0][2l+1], w[c1][2l], w[c1][2l+1]} and {w[c2][2l], w[c2][2l+1]}
  float4* wA = reinterpret_cast<float4*>(smem);                      // [49][32]
  float2* wB = reinterpret_cast<float2*>(smem + 49 * 32 * 4);        // [49][32]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float2* acc = reinterpret_cast<float2*>(smem + 147 * C0 + warp * kFT * kFT * C0) + lane;   // [64 outputs][32 lanes]
  for (int i = t; i < 49 * 32; i += blockDim.x) {
    const int tap = i >> 5, c = 2 * (i & 31);
    float w[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[k][0] = k < a.cin ? __ldg(a.w0 + (k * 49 + tap) * C0 + c) : 0.f;
      w[k][1] = k < a.cin ? __ldg(a.w0 + (k * 49 + tap) * C0 + c + 1) : 0.f;
    }
    wA[i] = make_float4(w[0][0], w[0][1], w[1][0], w[1][1]);
    wB[i] = make_float2(w[2][0], w[2][1]);
  }
  __syncthreads();
  wA += lane; wB += lane;
  const int ch = 2 * lane;
  const float2 bias = make_float2(__ldg(a.bias + ch), __ldg(a.bias + ch + 1));
  const __nv_bfloat162 zconst = bf2(bias);                 // an output no pixel reaches
  const float2 zc = __bfloat1622float2(zconst);
  double st1[2] = {0.0, 0.0}, st2[2] = {0.0, 0.0};         // this lane's two channels over all tiles of this warp

  const int tiles_x = (a.Ws + kFT - 1) / kFT, tiles_y = (a.Hs + kFT - 1) / kFT;
  const int per_image = tiles_x * tiles_y;
  const long long total = (long long)a.n_images * per_image;
  const size_t plane = (size_t)a.H * a.W;
  for (long long tile = (long long)blockIdx.x * kFWarps + warp; tile < total; tile += (long long)gridDim.x * kFWarps) {
    const int n = (int)(tile / per_image);
    const int rem = (int)(tile - (long long)n * per_image);
    const int tyi = rem / tiles_x;
    const int cy0 = tyi * kFT, cx0 = (rem - tyi * tiles_x) * kFT;
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;
    // ---- the 21 x 21 x cin window, lane = column: all loads issued before the first use
    const float* img = a.pixels + (size_t)n * a.cin * plane;
    float v[kFWin][3];
    const int x = ix0 + lane;
    const bool xok = lane < kFWin && x >= 0 && x < a.W;
#pragma unroll
    for (int yy = 0; yy < kFWin; ++yy) {
      const int y = iy0 + yy;
      const bool ok = xok && y >= 0 && y < a.H;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[yy][c] = (ok && c < a.cin) ? __ldg(img + c * plane + (size_t)y * a.W + x) : 0.f;
    }
    // ---- scatter, rows in order, columns in order (the order of the dense-window kernels: same bits)
    unsigned long long mask = 0ull;
#pragma unroll
    for (int yy = 0; yy < kFWin; ++yy) {
      const unsigned m = __ballot_sync(0xffffffffu, v[yy][0] != 0.f || v[yy][1] != 0.f || v[yy][2] != 0.f);
      if (m) mask = stem_scatter_row(m, v[yy][0], v[yy][1], v[yy][2], yy, mask, acc, wA, wB);
    }
    // ---- z0 = bf16(conv + bias) for every output of the tile, statistics of exactly the stored values
    float2 t1 = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
    int n_const = 0;
    __nv_bfloat16* zrow = a.z0 + (((size_t)n * a.Hs + cy0) * a.Ws + cx0) * C0 + ch;
#pragma unroll 1
    for (int cyl = 0; cyl < kFT; ++cyl) {
      if (cy0 + cyl >= a.Hs) break;
      const unsigned bits = (unsigned)(mask >> (cyl * kFT)) & 255u;
#pragma unroll
      for (int cxl = 0; cxl < kFT; ++cxl) {
        if (cx0 + cxl >= a.Ws) break;
        __nv_bfloat162 zb = zconst;
        if ((bits >> cxl) & 1u) {
          zb = bf2(__fadd2_rn(acc[(cyl * kFT + cxl) * 32], bias));
          const float2 z = __bfloat1622float2(zb);
          t1 = __fadd2_rn(t1, z);
          t2 = __ffma2_rn(z, z, t2);
        } else {
          ++n_const;
        }
        *reinterpret_cast<__nv_bfloat162*>(zrow + ((size_t)cyl * a.Ws + cxl) * C0) = zb;
      }
    }
    st1[0] += (double)t1.x + (double)n_const * (double)zc.x;
    st1[1] += (double)t1.y + (double)n_const * (double)zc.y;
    st2[0] += (double)t2.x + (double)n_const * ((double)zc.x * (double)zc.x);
    st2[1] += (double)t2.y + (double)n_const * ((double)zc.y * (double)zc.y);
    __syncwarp();
  }
  // per-CTA statistics: the warps of a channel are added in warp order
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem + 147 * C0);   // [warps][2][C0]
  red[(warp * 2) * C0 + ch] = st1[0]; red[(warp * 2) * C0 + ch + 1] = st1[1];
  red[(warp * 2 + 1) * C0 + ch] = st2[0]; red[(warp * 2 + 1) * C0 + ch + 1] = st2[1];
  __syncthreads();
  if (t < 2 * C0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kFWarps; ++k) s += red[(size_t)k * 2 * C0 + t];
    a.stat_parts[(size_t)blockIdx.x * 2 * C0 + t] = s;
  }
}

// All non-zero pixels of one 32-pixel chunk of an input row (ballot m), left to right: x[c] * dz[n, cy, cx, :] into this
// warp's accumulators.  Not inlined (the caller's chunk loop is unrolled: the row lives in registers).
__device__ __noinline__ void stem_wgrad_chunk(unsigned m, float r0, float r1, float r2, int x0, int y, int cin, int Hs, int Ws,
                                              const __nv_bfloat16* dz_img, float2* dw) {
  constexpr int C0 = 64;
  // conv rows this input row feeds: ky = y + 3 - 2 cy in [0, 6]
  const int cy_lo = y > 3 ? (y - 2) >> 1 : 0, cy_hi = min(Hs - 1, (y + 3) >> 1);
  while (m) {
    const int src = __ffs((int)m) - 1;
    m &= m - 1u;
    const int x = x0 + src;
    const float v0 = __shfl_sync(0xffffffffu, r0, src), v1 = __shfl_sync(0xffffffffu, r1, src), v2 = __shfl_sync(0xffffffffu, r2, src);
    const float2 vv[3] = {make_float2(v0, v0), make_float2(v1, v1), make_float2(v2, v2)};
    const int cx_lo = x > 3 ? (x - 2) >> 1 : 0, cx_hi = min(Ws - 1, (x + 3) >> 1);
    // the <= 4 x 4 gradient vectors this pixel meets, all loads in flight before the first use (the kernel is a chain of
    // L2 latencies otherwise: ncu long_scoreboard 6 cycles per issue with one row of loads at a time)
    float2 g[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cy = min(cy_lo + i, cy_hi);
      const __nv_bfloat16* grow = dz_img + ((size_t)cy * Ws) * C0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cx = min(cx_lo + j, cx_hi);
        g[i][j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(grow + (size_t)cx * C0));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (cy_lo + i > cy_hi) break;
      const int ky = y + 3 - 2 * (cy_lo + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cx_lo + j > cx_hi) break;
        const int tap = ky * 7 + (x + 3 - 2 * (cx_lo + j));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c >= cin) break;
          float2* d = dw + (c * 49 + tap) * 32;
          *d = __ffma2_rn(vv[c], g[i][j], *d);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kGWarps * 32, 1) stem_train_wgrad_kernel(const StemTrainArgs a) {
  constexpr int C0 = 64;
  extern __shared__ __align__(16) float smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int taps = a.cin * 49;
  float2* dw = reinterpret_cast<float2*>(smem + (size_t)warp * 147 * C0) + lane;   // [c * 49 + tap][32 lanes]
  for (int i = 0; i < taps; ++i) dw[i * 32] = make_float2(0.f, 0.f);
  const size_t plane = (size_t)a.H * a.W;
  const long long total = (long long)a.n_images * a.H;   // work unit: one input row of one image
  const int chunks = (a.W + 31) / 32;
  for (long long u = (long long)blockIdx.x * kGWarps + warp; u < total; u += (long long)gridDim.x * kGWarps) {
    const int n = (int)(u / a.H);
    const int y = (int)(u - (long long)n * a.H);
    const float* row = a.pixels + (size_t)n * a.cin * plane + (size_t)y * a.W;
    float v[kGChunks][3];
#pragma unroll
    for (int q = 0; q < kGChunks; ++q) {
      const int x = q * 32 + lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[q][c] = (q < chunks && x < a.W && c < a.cin) ? __ldg(row + c * plane + x) : 0.f;
    }
    const __nv_bfloat16* dz_img = a.dz + ((size_t)n * a.Hs * a.Ws) * C0 + 2 * lane;
#pragma unroll
    for (int q = 0; q < kGChunks; ++q) {
      const unsigned m = __ballot_sync(0xffffffffu, v[q][0] != 0.f || v[q][1] != 0.f || v[q][2] != 0.f);
      if (m) stem_wgrad_chunk(m, v[q][0], v[q][1], v[q][2], q * 32, y, a.cin, a.Hs, a.Ws, dz_img, dw);
    }
  }
  // the warps of the CTA, added in warp order -> the CTA's slot
  __syncthreads();
  float* part = a.dw_parts + (size_t)blockIdx.x * taps * C0;
  for (int i = t; i < taps * C0; i += blockDim.x) {
    const int r = i / C0, c = i - r * C0;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kGWarps; ++w) s += smem[(size_t)w * 147 * C0 + (r * 32 + (c >> 1)) * 2 + (c & 1)];
    part[i] = s;
  }
}

}  // namespace

// slots of the per-CTA partial results (MODE 0: doubles [slots][2][C0]; MODE 1: floats [slots][cin*49][C0])
int stem_train_slots(int n_images, int H, int W) {
  (void)H; (void)W;
  return n_images < 1 ? 1 : 148;   // upper bound of both kernels' grids (one slot per CTA)
}

static int stem_train_check(int cin, int C) {
  if (C != 64 || cin < 1 || cin > 3) return fail(TCVN_ERR_UNSUPPORTED, "training stem: built for <= 3 input and 64 output channels (got %d / %d)", cin, C);
  return TCVN_OK;
}

// z0 (bf16) = bias + conv0(pixels), stat_parts[slots][2][C] = per-CTA (sum, sum^2) of the stored z0 by channel
int stem_train_forward(const float* pixels, int n, int cin, int H, int W, const float* w0, const float* bias, int C, void* z0_bf16,
                       double* stat_parts, int* n_slots, cudaStream_t stream) {
  TCVN_TRY(stem_train_check(cin, C));
  StemTrainArgs a{};
  a.pixels = pixels; a.n_images = n; a.cin = cin; a.H = H; a.W = W;
  a.Hs = (H + 6 - 7) / 2 + 1; a.Ws = (W + 6 - 7) / 2 + 1;
  a.w0 = w0; a.bias = bias; a.z0 = static_cast<__nv_bfloat16*>(z0_bf16); a.stat_parts = stat_parts;
  const long long tiles = (long long)n * ceil_div(a.Hs, kFT) * ceil_div(a.Ws, kFT);
  const long long want = ceil_div_ll(tiles, kFWarps);
  const int grid = (int)(want < 148 ? (want < 1 ? 1 : want) : 148);
  *n_slots = grid;
  const size_t smem = ((size_t)147 * 64 + (size_t)kFWarps * kFT * kFT * 64) * 4;
  TCVN_CUDA(cudaFuncSetAttribute(stem_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stem_train_fwd_kernel<<<grid, kFWarps * 32, smem, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// dw_parts[slots][cin*49][C] = per-CTA partial sums of conv0's weight gradient (dz bf16 [n, Hs, Ws, C])
int stem_train_wgrad(const float* pixels, int n, int cin, int H, int W, const void* dz_bf16, int C, float* dw_parts, int* n_slots,
                     cudaStream_t stream) {
  TCVN_TRY(stem_train_check(cin, C));
  StemTrainArgs a{};
  a.pixels = pixels; a.n_images = n; a.cin = cin; a.H = H; a.W = W;
  a.Hs = (H + 6 - 7) / 2 + 1; a.Ws = (W + 6 - 7) / 2 + 1;
  a.dz = static_cast<const __nv_bfloat16*>(dz_bf16); a.dw_parts = dw_parts;
  if (W > 32 * kGChunks) return fail(TCVN_ERR_UNSUPPORTED, "training stem: image width %d (weight-gradient kernel holds rows of <= %d pixels)", W, 32 * kGChunks);
  const long long want = ceil_div_ll((long long)n * H, kGWarps);
  const int grid = (int)(want < 148 ? (want < 1 ? 1 : want) : 148);
  *n_slots = grid;
  const size_t smem = (size_t)kGWarps * 147 * 64 * 4;
  TCVN_CUDA(cudaFuncSetAttribute(stem_train_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stem_train_wgrad_kernel<<<grid, kGWarps * 32, smem, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn
