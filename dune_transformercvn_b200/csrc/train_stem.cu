// Stem of the TRAINING path (dense_net.py:112-118 under autograd): raw conv0 output + its batch statistics, and conv0's
// weight gradient.  Both are hit-driven like the inference stem (simt.cu) and both are bit-reproducible: a persistent CTA
// walks 16 x 16 tiles of conv outputs in a fixed order, the non-zero pixels of the tile's input window are compacted in
// a fixed order (ballot + prefix sums), every accumulator has exactly one owner thread, and the per-CTA partial results
// land in fixed slots that a later kernel adds in a fixed order.  (Round 1 scattered hits with float atomics: the one-ulp
// run-to-run differences of z0 were amplified by the 67 train-mode BatchNorms to 20 % of the gradient.)
//
//   MODE 0  z0[n, cy, cx, :] = bf16(b + conv7x7s2p3(pixels))   and  parts[cta][2][C0] = per-CTA (sum z, sum z^2) in double
//   MODE 1  parts[cta][cin*49][C0] = sum over the CTA's tiles of  x[2cy-3+ky, 2cx-3+kx, c] * dz[n, cy, cx, :]
#include "kernels.h"
#include "stem.cuh"

namespace tcvn {

namespace {

constexpr int kTileC = 16;   // conv outputs per tile edge that the tile OWNS (it evaluates 17: the stem_scatter geometry)

struct StemTrainArgs {
  const float* pixels; int n_images, cin, H, W, Hs, Ws;
  const float* w0;        // [cin*49][C0]           (MODE 0)
  const float* bias;      // [C0]                   (MODE 0)
  __nv_bfloat16* z0;      // [n, Hs, Ws, C0] bf16   (MODE 0)
  double* stat_parts;     // [grid][2][C0]          (MODE 0)
  const __nv_bfloat16* dz; // [n, Hs, Ws, C0] bf16  (MODE 1)
  float* dw_parts;        // [grid][cin*49][C0]     (MODE 1)
};

template <int MODE, int C0>
__global__ void __launch_bounds__(kStemThreads, 1) stem_train_kernel(const StemTrainArgs a) {
  extern __shared__ __align__(16) float smem[];
  // MODE 0: filter bank [cin*49][C0] | accumulators [289][C0] | hits      MODE 1: dz tile [256][C0] bf16 | hits
  float* wsm = smem;
  float* acc = MODE == 0 ? wsm + a.cin * 49 * C0 : smem;
  __nv_bfloat16* dzs = reinterpret_cast<__nv_bfloat16*>(smem);
  float4* hits = MODE == 0 ? reinterpret_cast<float4*>(acc + kStemTC * kStemTC * C0)
                           : reinterpret_cast<float4*>(smem + kTileC * kTileC * C0 / 2);
  __shared__ int row_count[kStemRows];
  __shared__ int row_start[kStemRows + 1];
  __shared__ int touched[kStemTC * kStemTC];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int ch = t & (C0 - 1);
  if (MODE == 0)
    for (int i = t; i < a.cin * 49 * C0; i += blockDim.x) wsm[i] = __ldg(a.w0 + i);
  const int tiles_x = (a.Ws + kTileC - 1) / kTileC, tiles_y = (a.Hs + kTileC - 1) / kTileC;
  const int per_image = tiles_x * tiles_y;
  const long long total = (long long)a.n_images * per_image;
  const size_t plane = (size_t)a.H * a.W;
  // the window of the NEXT tile is loaded into registers while this tile is processed (cf. stem_fused_kernel)
  float v[3][2][3];
  auto load_window = [&](long long tile_id) {
    const int n = (int)(tile_id / per_image);
    const int rem = (int)(tile_id - (long long)n * per_image);
    const int iy0 = 2 * (rem / tiles_x) * kTileC - 3, ix0 = 2 * (rem % tiles_x) * kTileC - 3;
    const float* img = a.pixels + (size_t)n * a.cin * plane;
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs, y = iy0 + yy;
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        const int xx = lane + 32 * cs, x = ix0 + xx;
        const bool ok = yy < kStemIn && xx < kStemIn && y >= 0 && y < a.H && x >= 0 && x < a.W;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[rs][cs][c] = (ok && c < a.cin) ? __ldg(img + c * plane + (size_t)y * a.W + x) : 0.f;
      }
    }
  };
  // MODE 0 state: statistics of this thread's channel over the positions it writes
  double st1 = 0.0, st2 = 0.0;
  const float bias = MODE == 0 ? __ldg(a.bias + ch) : 0.f;
  // MODE 1 state: this thread's filter taps tg, tg + 8, ... (7 or 6 of the 49) x 3 input channels
  const int tg = t >> 6;
  float dwacc[7][3];
  int tap_ky[7], tap_kx[7];   // ky = 99 marks the missing seventh tap
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    dwacc[q][0] = 0.f; dwacc[q][1] = 0.f; dwacc[q][2] = 0.f;
    const int tap = tg + 8 * q;
    tap_ky[q] = tap < 49 ? tap / 7 : 99;
    tap_kx[q] = tap % 7;
  }

  if ((long long)blockIdx.x < total) load_window(blockIdx.x);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = (int)(tile / per_image);
    const int rem = (int)(tile - (long long)n * per_image);
    const int cy0 = (rem / tiles_x) * kTileC, cx0 = (rem % tiles_x) * kTileC;
    __syncthreads();  // previous tile fully consumed
    // ---- compact the non-zero pixels of the window: rows in order, columns in order
    unsigned m0[3], m1[3];
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs;
      m0[rs] = __ballot_sync(0xffffffffu, v[rs][0][0] != 0.f || v[rs][0][1] != 0.f || v[rs][0][2] != 0.f);
      m1[rs] = __ballot_sync(0xffffffffu, v[rs][1][0] != 0.f || v[rs][1][1] != 0.f || v[rs][1][2] != 0.f);
      if (lane == 0 && yy < kStemIn) row_count[yy] = __popc(m0[rs]) + __popc(m1[rs]);
    }
    if (MODE == 0) {
      float4* a4 = reinterpret_cast<float4*>(acc);
      for (int i = t; i < kStemTC * kStemTC * C0 / 4; i += blockDim.x) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < kStemTC * kStemTC) touched[t] = 0;
    }
    __syncthreads();
    if (warp == 0) {
      int run = 0;
      for (int base = 0; base < kStemRows; base += 32) {
        const int r = base + lane;
        const int c = r < kStemRows ? row_count[r] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        if (r < kStemRows) row_start[r] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) row_start[kStemRows] = run;
    }
    __syncthreads();
    const int nhits = row_start[kStemRows];
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs;
      if (yy < kStemIn) {
        const int base = row_start[yy];
        const unsigned below = (1u << lane) - 1u;
        if (m0[rs] >> lane & 1u)
          hits[base + __popc(m0[rs] & below)] = make_float4(__int_as_float(yy * 64 + lane), v[rs][0][0], v[rs][0][1], v[rs][0][2]);
        if (m1[rs] >> lane & 1u)
          hits[base + __popc(m0[rs]) + __popc(m1[rs] & below)] =
              make_float4(__int_as_float(yy * 64 + lane + 32), v[rs][1][0], v[rs][1][1], v[rs][1][2]);
      }
    }
    if (tile + gridDim.x < total) load_window(tile + gridDim.x);  // in flight during the rest of this tile
    if (MODE == 1 && nhits > 0) {
      // gradient tile of the owned conv outputs (zero outside the map), 16-byte loads
      constexpr int VPR = C0 / 8;   // uint4 per position
      for (int i = t; i < kTileC * kTileC * VPR; i += blockDim.x) {
        const int p = i / VPR, q = i - p * VPR;
        const int cy = cy0 + (p >> 4), cx = cx0 + (p & 15);
        uint4 g = make_uint4(0u, 0u, 0u, 0u);
        if (cy < a.Hs && cx < a.Ws)
          g = __ldg(reinterpret_cast<const uint4*>(a.dz + (((size_t)n * a.Hs + cy) * a.Ws + cx) * C0) + q);
        reinterpret_cast<uint4*>(dzs)[i] = g;
      }
    }
    __syncthreads();
    if (MODE == 0) {
      stem_scatter<C0, false>(wsm, acc, hits, nhits, touched, a.cin, ch, t >> 7, (t >> 6) & 1);
      __syncthreads();
      // owned outputs -> z0 (+ bias); 8 positions x C0 channels per step, 256 bytes per position
      for (int p = t >> 6; p < kTileC * kTileC; p += kStemThreads / C0) {
        const int cyl = p >> 4, cxl = p & 15;
        const int cy = cy0 + cyl, cx = cx0 + cxl;
        if (cy < a.Hs && cx < a.Ws) {
          // stored (and counted in the statistics) as bf16: three more passes read this map (pooling, BN0 backward x 2)
          const __nv_bfloat16 zb = __float2bfloat16_rn(acc[(cyl * kStemTC + cxl) * C0 + ch] + bias);
          a.z0[(((size_t)n * a.Hs + cy) * a.Ws + cx) * C0 + ch] = zb;
          const float z = __bfloat162float(zb);
          st1 += (double)z;
          st2 += (double)z * (double)z;
        }
      }
    } else {
      for (int h = 0; h < nhits; ++h) {
        const float4 hit = hits[h];
        const int code = __float_as_int(hit.x);
        const int yy = code >> 6, xx = code & 63;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int dy = yy - tap_ky[q], dx = xx - tap_kx[q];
          if (dy >= 0 && dx >= 0 && !((dy | dx) & 1) && dy < 2 * kTileC && dx < 2 * kTileC) {
            const float g = __bfloat162float(dzs[((dy >> 1) * kTileC + (dx >> 1)) * C0 + ch]);
            dwacc[q][0] = fmaf(hit.y, g, dwacc[q][0]);
            dwacc[q][1] = fmaf(hit.z, g, dwacc[q][1]);
            dwacc[q][2] = fmaf(hit.w, g, dwacc[q][2]);
          }
        }
      }
    }
  }
  if (MODE == 0) {
    // per-CTA statistics: the 8 position lanes of a channel are added in a fixed order
    __syncthreads();
    double* red = reinterpret_cast<double*>(acc);   // [8][C0][2]
    red[((t >> 6) * C0 + ch) * 2] = st1;
    red[((t >> 6) * C0 + ch) * 2 + 1] = st2;
    __syncthreads();
    if (t < C0) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < kStemThreads / C0; ++k) { s1 += red[(k * C0 + t) * 2]; s2 += red[(k * C0 + t) * 2 + 1]; }
      a.stat_parts[((size_t)blockIdx.x * 2) * C0 + t] = s1;
      a.stat_parts[((size_t)blockIdx.x * 2 + 1) * C0 + t] = s2;
    }
  } else {
    float* part = a.dw_parts + (size_t)blockIdx.x * a.cin * 49 * C0;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const int tap = tg + 8 * q;
      if (tap < 49) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < a.cin) part[(c * 49 + tap) * C0 + ch] = dwacc[q][c];
      }
    }
  }
}

}  // namespace

// slots of the per-CTA partial results (MODE 0: doubles [slots][2][C0]; MODE 1: floats [slots][cin*49][C0])
int stem_train_slots(int n_images, int H, int W) {
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  const long long tiles = (long long)n_images * ceil_div(Hs, kTileC) * ceil_div(Ws, kTileC);
  return (int)(tiles < 148 ? (tiles < 1 ? 1 : tiles) : 148);
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
  const int grid = stem_train_slots(n, H, W);
  *n_slots = grid;
  const size_t smem = ((size_t)cin * 49 * 64 + (size_t)kStemTC * kStemTC * 64) * 4 + (size_t)kStemIn * kStemIn * 16;
  TCVN_CUDA(cudaFuncSetAttribute(stem_train_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stem_train_kernel<0, 64><<<grid, kStemThreads, smem, stream>>>(a);
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
  const int grid = stem_train_slots(n, H, W);
  *n_slots = grid;
  const size_t smem = (size_t)kTileC * kTileC * 64 * 2 + (size_t)kStemIn * kStemIn * 16;
  TCVN_CUDA(cudaFuncSetAttribute(stem_train_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stem_train_kernel<1, 64><<<grid, kStemThreads, smem, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn
