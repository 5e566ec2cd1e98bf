// Hit-driven stem of the inference path, one WARP per tile (dense_net.py:111-122 fed from the COO hit list of
// fully_sparse_dataset.py / neutrino_full_dense_trainer.py:15-24,59-60).
//
// The round-1 kernel (simt.cu: a 512-thread CTA per 8 x 8 pooled tile) spent its time on per-tile fixed costs: four CTA
// barriers, zeroing 74 KB of accumulators, a classification pass run by one warp while fifteen waited - for ~3 hits per
// tile (the pixel maps are 0.2-1 % occupied).  Here a tile is 4 x 4 pooled pixels (9 x 9 conv outputs <- 23 x 23 input
// pixels) and belongs to ONE warp: lane l owns channels 2l, 2l+1 of the tile's accumulators (20 KB of shared memory per
// warp), every lane sees the same hits, so all control flow is warp-uniform and there is no barrier at all.  The
// accumulators are never zeroed: a bit mask of touched conv outputs (81 bits in registers) tells first touches (start
// from 0) from later ones (read-modify-write), which conv outputs have to be activated, and which pooled pixels are the
// per-channel constant.  Same arithmetic and the same summation order per conv output (hit-list order) as the dense
// window kernel stem_fused_kernel: the two stay bit-identical.
//
// Binning (once per forward, all images): hits -> per-tile record lists in hit-list order.  count: shared-memory integer
// atomics per image; fill: one warp per image takes 8 hits x 4 candidate tiles per step, __match_any_sync gives every
// (hit, tile) pair its rank among the pairs of the same tile in hit order.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "umma.h"
#include "stem.cuh"

namespace tcvn {
namespace {

constexpr int kWP = 4;                  // pooled tile edge
constexpr int kWC = 2 * kWP + 1;        // 9 conv outputs per edge
constexpr int kWIn = 2 * kWC + 5;       // 23 input pixels per edge
constexpr int kWStep = 4 * kWP;         // 16 input pixels between tile origins
constexpr int kWPos = kWC * kWC;        // 81 conv outputs per tile
constexpr int kWarps = 9;               // warps (= tiles in flight) per CTA
constexpr int kC0 = 64;
constexpr int kTaps = 147;              // 3 x 7 x 7 (rows of missing input channels are zero)

__host__ __device__ inline int tiles_of(int pooled) { return (pooled + kWP - 1) / kWP; }

// tiles [lo, hi] whose input window [16 t - 3, 16 t + 19] holds coordinate v
__device__ __forceinline__ void tile_range(int v, int tiles, int& lo, int& hi) {
  lo = v >= kWIn - 3 ? (v - 4) >> 4 : 0;
  hi = min(tiles - 1, (v + 3) >> 4);
}

// ---------------------------------------------------------------------------------------------- binning
// grid = images; counts the records of every tile of the image, leaves the image-local exclusive scan in
// tile_local[n][per_image] and the image's record count in image_total[n]
__global__ void __launch_bounds__(256) stem_bin_count_kernel(const int32_t* __restrict__ coords,
                                                             const long long* __restrict__ image_offsets, int H, int W,
                                                             int tiles_x, int tiles_y, int32_t* __restrict__ tile_local,
                                                             int32_t* __restrict__ image_total) {
  extern __shared__ int cnt[];           // [per_image]
  __shared__ int warp_sum[8];
  const int n = blockIdx.x, t = threadIdx.x;
  const int per_image = tiles_x * tiles_y;
  for (int i = t; i < per_image; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  const long long lo = __ldg(image_offsets + n), hi = __ldg(image_offsets + n + 1);
  for (long long h = lo + t; h < hi; h += blockDim.x) {
    const int y = __ldg(coords + 3 * h + 1), x = __ldg(coords + 3 * h + 2);
    if (y < 0 || y >= H || x < 0 || x >= W) continue;   // out-of-map hits are dropped (sparse_to_dense would raise)
    int ty0, ty1, tx0, tx1;
    tile_range(y, tiles_y, ty0, ty1);
    tile_range(x, tiles_x, tx0, tx1);
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&cnt[ty * tiles_x + tx], 1);
  }
  __syncthreads();
  // exclusive scan of cnt[per_image]: each thread owns a run of consecutive tiles
  const int per = (per_image + blockDim.x - 1) / blockDim.x;
  const int b = t * per, e = min(per_image, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += cnt[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if ((t & 31) >= o) inc += v;
  }
  if ((t & 31) == 31) warp_sum[t >> 5] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (t >> 5); ++w) base += warp_sum[w];
  int run = base + inc - s;
  for (int i = b; i < e; ++i) { tile_local[(size_t)n * per_image + i] = run; run += cnt[i]; }
  if (t == blockDim.x - 1) image_total[n] = base + inc;
}

// exclusive scan of n int32 by one block: start[i], start[n] = total
__global__ void __launch_bounds__(1024) stem_scan_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ start) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = t * per, e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - s;
  for (int i = b; i < e; ++i) { start[i] = run; run += counts[i]; }
  if (t == 1023) start[n] = part[1023];
}

// one warp per image: writes tile_start (absolute, [n_images * per_image + 1]) and the records in hit-list order.
// A step takes 8 hits x 4 candidate tiles (lane = 4 * hit + candidate); lanes are ordered by hit, so the rank of a lane
// among the lanes with the same tile IS the hit order.  Hits are loaded 32 at a time, one chunk ahead.
template <typename V>
__global__ void __launch_bounds__(32) stem_bin_fill_kernel(const int32_t* __restrict__ coords, const V* __restrict__ values,
                                                           const long long* __restrict__ image_offsets, int n_images, int cin,
                                                           float divisor, int H, int W, int tiles_x, int tiles_y,
                                                           const int32_t* __restrict__ tile_local,
                                                           const int32_t* __restrict__ image_base,
                                                           int32_t* __restrict__ tile_start, float4* __restrict__ recs) {
  extern __shared__ int cur[];           // [per_image] next free record slot of every tile
  const int n = blockIdx.x, lane = threadIdx.x;
  const int per_image = tiles_x * tiles_y;
  const int base = __ldg(image_base + n);
  for (int i = lane; i < per_image; i += 32) {
    const int s = base + __ldg(tile_local + (size_t)n * per_image + i);
    cur[i] = s;
    tile_start[(size_t)n * per_image + i] = s;
  }
  if (n == n_images - 1 && lane == 0) tile_start[(size_t)n_images * per_image] = __ldg(image_base + n_images);
  __syncwarp();
  const long long lo = __ldg(image_offsets + n), hi = __ldg(image_offsets + n + 1);
  const int sub = lane >> 2, dy = (lane >> 1) & 1, dx = lane & 1;
  auto load = [&](long long h, int& y, int& x, float& v0, float& v1, float& v2) {
    y = -1; x = -1; v0 = v1 = v2 = 0.f;
    if (h < hi) {
      y = __ldg(coords + 3 * h + 1); x = __ldg(coords + 3 * h + 2);
      float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c < cin) {
          v[c] = static_cast<float>(values[h * cin + c]);
          if (divisor != 0.f) v[c] = __fdiv_rn(v[c], divisor);   // same bits as the reference's v / 255.0
        }
      v0 = v[0]; v1 = v[1]; v2 = v[2];
    }
  };
  int ny, nx; float n0, n1, n2;
  load(lo + lane, ny, nx, n0, n1, n2);
  for (long long h0 = lo; h0 < hi; h0 += 32) {
    const int cy = ny, cx = nx; const float c0 = n0, c1 = n1, c2 = n2;
    load(h0 + 32 + lane, ny, nx, n0, n1, n2);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int src = 8 * s + sub;
      const int y = __shfl_sync(0xffffffffu, cy, src), x = __shfl_sync(0xffffffffu, cx, src);
      const float v0 = __shfl_sync(0xffffffffu, c0, src), v1 = __shfl_sync(0xffffffffu, c1, src),
                  v2 = __shfl_sync(0xffffffffu, c2, src);
      bool valid = y >= 0 && y < H && x >= 0 && x < W;
      int key = 0x40000000 | lane, ty = 0, tx = 0;
      if (valid) {
        int ty0, ty1, tx0, tx1;
        tile_range(y, tiles_y, ty0, ty1);
        tile_range(x, tiles_x, tx0, tx1);
        ty = ty0 + dy; tx = tx0 + dx;
        valid = ty <= ty1 && tx <= tx1;
        if (valid) key = ty * tiles_x + tx;
      }
      const unsigned m = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(m) - 1;
      int slot = 0;
      if (valid && lane == leader) { slot = cur[key]; cur[key] = slot + __popc(m); }
      slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(m & ((1u << lane) - 1u));
      if (valid) {
        const int yy = y - (kWStep * ty - 3), xx = x - (kWStep * tx - 3);
        recs[slot] = make_float4(__int_as_float(yy * 64 + xx), v0, v1, v2);
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------- the tile kernel
template <typename TO> struct Pair;
template <> struct Pair<float> {
  using T = float2;
  static __device__ __forceinline__ T make(float2 a) { return a; }
};
template <> struct Pair<__nv_bfloat16> {
  using T = __nv_bfloat162;
  static __device__ __forceinline__ T make(float2 a) { return __float22bfloat162_rn(a); }
};

__device__ __forceinline__ float2 prelu2(float2 y, float a0, float a1) {
  return make_float2(prelu(y.x, a0), prelu(y.y, a1));
}

// touched conv outputs of a tile: bit cy * 9 + cx; rows 0..6 in lo, rows 7..8 in hi
struct Mask81 {
  unsigned long long lo = 0ull;
  unsigned hi = 0u;
  template <int R> __device__ __forceinline__ unsigned row() const {   // 9 bits of conv row R (compile-time)
    return R < 7 ? (unsigned)(lo >> (9 * R)) & 511u : (hi >> (9 * (R - 7))) & 511u;
  }
};

template <typename TO>
__global__ void __launch_bounds__(kWarps * 32, 1) stem_warp_kernel(const int32_t* __restrict__ tile_start,
                                                                   const float4* __restrict__ recs, int image0,
                                                                   int n_images, int cin, const float* __restrict__ w0,
                                                                   const float* __restrict__ s_scale,
                                                                   const float* __restrict__ s_shift,
                                                                   const float* __restrict__ s_alpha, TO* __restrict__ blk,
                                                                   int ldo, int Hb, int Wb) {
  extern __shared__ __align__(16) float smem[];
  // filter bank re-laid out per (tap, lane): {w[c0][2l], w[c0][2l+1], w[c1][2l], w[c1][2l+1]} and {w[c2][2l], w[c2][2l+1]};
  // absent input channels are zero
  float4* wA = reinterpret_cast<float4*>(smem);                      // [49][32]
  float2* wB = reinterpret_cast<float2*>(smem + 49 * 32 * 4);        // [49][32]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float2* acc = reinterpret_cast<float2*>(smem + kTaps * kC0 + warp * kWPos * kC0) + lane;   // [81][32 lanes] float2
  for (int i = t; i < 49 * 32; i += blockDim.x) {
    const int tap = i >> 5, c = 2 * (i & 31);
    float w[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[k][0] = k < cin ? __ldg(w0 + (k * 49 + tap) * kC0 + c) : 0.f;
      w[k][1] = k < cin ? __ldg(w0 + (k * 49 + tap) * kC0 + c + 1) : 0.f;
    }
    wA[i] = make_float4(w[0][0], w[0][1], w[1][0], w[1][1]);
    wB[i] = make_float2(w[2][0], w[2][1]);
  }
  __syncthreads();
  wA += lane; wB += lane;

  const int ch = 2 * lane;
  const float2 sc = make_float2(__ldg(s_scale + ch), __ldg(s_scale + ch + 1));
  const float2 sh = make_float2(__ldg(s_shift + ch), __ldg(s_shift + ch + 1));
  const float al0 = __ldg(s_alpha + ch), al1 = __ldg(s_alpha + ch + 1);
  // a conv output no hit reached is the per-channel constant PReLU(shift); nine of them pool to a constant
  // (evaluated with the same operations as the general case, so the shortcut is bit-identical)
  const float2 ca = prelu2(make_float2(fmaf(0.f, sc.x, sh.x), fmaf(0.f, sc.y, sh.y)), al0, al1);
  float2 cp = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 9; ++k) cp = __fadd2_rn(cp, ca);
  using P2 = typename Pair<TO>::T;
  const P2 cpool = Pair<TO>::make(pool_avg9<TO>(cp));

  const int tiles_x = tiles_of(Wb), tiles_y = tiles_of(Hb);
  const int per_image = tiles_x * tiles_y;
  const int pix_pitch = ldo, row_pitch = (Wb + 2) * ldo;   // elements between pooled pixels / pooled rows
  const long long first = (long long)image0 * per_image;
  const long long total = (long long)n_images * per_image;
  const int nwarps = blockDim.x >> 5;
  const long long stride = (long long)gridDim.x * nwarps;
  long long tile = (long long)blockIdx.x * nwarps + warp;
  if (tile >= total) return;

  // bounds two tiles ahead, records one tile ahead
  int b0 = __ldg(tile_start + first + tile), e0 = __ldg(tile_start + first + tile + 1);
  int b1 = 0, e1 = 0;
  if (tile + stride < total) { b1 = __ldg(tile_start + first + tile + stride); e1 = __ldg(tile_start + first + tile + stride + 1); }
  float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < e0 - b0) rec = __ldg(recs + b0 + lane);

  for (; tile < total; tile += stride) {
    const int n = (int)((unsigned)tile / (unsigned)per_image);   // launch_stem_bin keeps the tile count below 2^31
    const int rem = (int)tile - n * per_image;
    const int tyi = rem / tiles_x;
    const int py0 = tyi * kWP, px0 = (rem - tyi * tiles_x) * kWP;
    const int cnt = e0 - b0;
    int b2 = 0, e2 = 0;
    if (tile + 2 * stride < total) {
      b2 = __ldg(tile_start + first + tile + 2 * stride); e2 = __ldg(tile_start + first + tile + 2 * stride + 1);
    }
    float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < e1 - b1) rec_next = __ldg(recs + b1 + lane);
    P2* out = reinterpret_cast<P2*>(blk + ((size_t)n * (Hb + 2) * (Wb + 2) + (size_t)(py0 + 1) * (Wb + 2) + (px0 + 1)) * ldo + ch);

    if (cnt == 0) {   // nothing reaches this tile: the constant vector everywhere
#pragma unroll
      for (int pyl = 0; pyl < kWP; ++pyl)
#pragma unroll
        for (int pxl = 0; pxl < kWP; ++pxl)
          if (py0 + pyl < Hb && px0 + pxl < Wb)
            *reinterpret_cast<P2*>(reinterpret_cast<TO*>(out) + pyl * row_pitch + pxl * pix_pitch) = cpool;
    } else {
      // ---- (1) scatter: every hit adds v * w[tap] to the <= 4 x 4 conv outputs it reaches.  Four outputs of a conv row
      // are in flight together; a slot beyond the row's last output repeats a legal address and is not stored.
      Mask81 M;
      for (int r0 = 0; r0 < cnt; r0 += 32) {
        if (r0 > 0) {   // dense windows only: further records of this tile
          rec = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r0 + lane < cnt) rec = __ldg(recs + b0 + r0 + lane);
        }
        const int m = min(32, cnt - r0);
        for (int h = 0; h < m; ++h) {
          const int code = __float_as_int(__shfl_sync(0xffffffffu, rec.x, h));
          const float v0 = __shfl_sync(0xffffffffu, rec.y, h), v1 = __shfl_sync(0xffffffffu, rec.z, h),
                      v2 = __shfl_sync(0xffffffffu, rec.w, h);
          const float2 vv0 = make_float2(v0, v0), vv1 = make_float2(v1, v1), vv2 = make_float2(v2, v2);
          const int yy = code >> 6, xx = code & 63;
          const int cy_lo = yy > 6 ? (yy - 5) >> 1 : 0, cy_hi = min(kWC - 1, yy >> 1);
          const int cx_lo = xx > 6 ? (xx - 5) >> 1 : 0, cx_hi = min(kWC - 1, xx >> 1);
          const int ncx = cx_hi - cx_lo + 1;                 // 1..4
          const unsigned run = (1u << ncx) - 1u;
          for (int cy = cy_lo; cy <= cy_hi; ++cy) {
            const int pos0 = cy * kWC + cx_lo;
            const int tap0 = (yy - 2 * cy) * 7 + xx - 2 * cx_lo;
            const bool low = cy < 7;
            const int sft = low ? pos0 : pos0 - 63;
            const unsigned field = low ? (unsigned)(M.lo >> sft) : M.hi >> sft;   // bit j: output j of the run was touched
            float2 a[4];
            float4 wa[4];
            float2 wb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int jj = min(j, ncx - 1);
              wa[j] = wA[(tap0 - 2 * jj) * 32];
              wb[j] = wB[(tap0 - 2 * jj) * 32];
              a[j] = make_float2(0.f, 0.f);
              if ((field >> jj) & 1u) a[j] = acc[(pos0 + jj) * 32];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a[j] = __ffma2_rn(vv0, make_float2(wa[j].x, wa[j].y), a[j]);
              a[j] = __ffma2_rn(vv1, make_float2(wa[j].z, wa[j].w), a[j]);
              a[j] = __ffma2_rn(vv2, wb[j], a[j]);
              if (j < ncx) acc[(pos0 + j) * 32] = a[j];
            }
            if (low) M.lo |= (unsigned long long)run << sft; else M.hi |= run << sft;
          }
        }
      }

      // ---- (2) bias + BN0 + PReLU0 in place on the outputs that were reached (the set bits of a row, four per step);
      // the others get the constant, so that (3) reads blindly
#define TCVN_STEM_ROW(R)                                                                  \
      {                                                                                   \
        const unsigned row_bits = M.row<R>();                                             \
        float2* rowp = acc + (R) * kWC * 32;                                              \
        if (row_bits == 0u) {                                                             \
          _Pragma("unroll") for (int cx = 0; cx < kWC; ++cx) rowp[cx * 32] = ca;          \
        } else {                                                                          \
          unsigned bits = row_bits;                                                       \
          while (bits) {                                                                  \
            int c[4];                                                                     \
            float2 v[4];                                                                  \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) {                               \
              c[j] = __ffs((int)bits) - 1;                                                \
              bits &= bits - 1u;                                                          \
              v[j] = rowp[max(c[j], 0) * 32];                                             \
            }                                                                             \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) {                               \
              v[j] = prelu2(__ffma2_rn(v[j], sc, sh), al0, al1);                          \
              if (c[j] >= 0) rowp[c[j] * 32] = v[j];                                      \
            }                                                                             \
          }                                                                               \
          _Pragma("unroll") for (int cx = 0; cx < kWC; ++cx)                              \
            if (!((row_bits >> cx) & 1u)) rowp[cx * 32] = ca;                             \
        }                                                                                 \
      }
      TCVN_STEM_ROW(0) TCVN_STEM_ROW(1) TCVN_STEM_ROW(2) TCVN_STEM_ROW(3) TCVN_STEM_ROW(4)
      TCVN_STEM_ROW(5) TCVN_STEM_ROW(6) TCVN_STEM_ROW(7) TCVN_STEM_ROW(8)
#undef TCVN_STEM_ROW

      // ---- (3) AvgPool2d(3, 2) -> channels [0, 64) of the ringed block buffer
#define TCVN_STEM_POOLROW(PYL)                                                                                 \
      if (py0 + (PYL) < Hb) {                                                                                  \
        const unsigned wrow = M.row<2 * (PYL)>() | M.row<2 * (PYL) + 1>() | M.row<2 * (PYL) + 2>();            \
        _Pragma("unroll") for (int pxl = 0; pxl < kWP; ++pxl) {                                                \
          if (px0 + pxl >= Wb) continue;                                                                       \
          P2* dst = reinterpret_cast<P2*>(reinterpret_cast<TO*>(out) + (PYL) * row_pitch + pxl * pix_pitch);   \
          if ((wrow & (7u << (2 * pxl))) == 0u) {                                                              \
            *dst = cpool;                                                                                      \
          } else {                                                                                             \
            float2 s = make_float2(0.f, 0.f);                                                                  \
            _Pragma("unroll") for (int d = 0; d < 9; ++d)                                                      \
              s = __fadd2_rn(s, acc[((2 * (PYL) + d / 3) * kWC + 2 * pxl + d % 3) * 32]);                      \
            *dst = Pair<TO>::make(pool_avg9<TO>(s));                                                           \
          }                                                                                                    \
        }                                                                                                      \
      }
      TCVN_STEM_POOLROW(0) TCVN_STEM_POOLROW(1) TCVN_STEM_POOLROW(2) TCVN_STEM_POOLROW(3)
#undef TCVN_STEM_POOLROW
    }
    b0 = b1; e0 = e1; b1 = b2; e1 = e2;
    rec = rec_next;
  }
}

struct BinLayout {
  size_t local, totals, base, start, recs, bytes;
};
BinLayout bin_layout(long long n_images, int Hb, int Wb, long long nnz) {
  const size_t tiles = (size_t)n_images * tiles_of(Hb) * tiles_of(Wb);
  BinLayout L;
  size_t off = 0;
  L.local = off;  off += align_up(tiles * sizeof(int32_t), 256);
  L.totals = off; off += align_up((size_t)(n_images + 1) * sizeof(int32_t), 256);
  L.base = off;   off += align_up((size_t)(n_images + 1) * sizeof(int32_t), 256);
  L.start = off;  off += align_up((tiles + 1) * sizeof(int32_t), 256);
  L.recs = off;   off += align_up((size_t)(4 * nnz + 4) * sizeof(float4), 256);
  L.bytes = off;
  return L;
}

}  // namespace

size_t stem_bins_bytes(int n_images, int Hb, int Wb, long long nnz) { return bin_layout(n_images, Hb, Wb, nnz).bytes; }

// bins the hits of ALL n_images of a forward call (image_offsets[i] = first hit of image i); bins = stem_bins_bytes(...)
int launch_stem_bin(const int32_t* coords, const void* values, bool values_u8, const long long* image_offsets, int n_images,
                    long long nnz, int cin, float divisor, int H, int W, int Hb, int Wb, void* bins, cudaStream_t stream) {
  if (n_images == 0) return TCVN_OK;
  if (cin > 3) return fail(TCVN_ERR_UNSUPPORTED, "stem: %d input channels (kernel handles up to 3)", cin);
  const int tiles_x = tiles_of(Wb), tiles_y = tiles_of(Hb);
  const int per_image = tiles_x * tiles_y;
  if ((long long)n_images * per_image >= (1ll << 31) - 2 || 4 * nnz >= (1ll << 31) - 8)
    return fail(TCVN_ERR_UNSUPPORTED, "stem: too many tiles / hits in one call");
  if ((size_t)per_image * sizeof(int) > 96 * 1024) return fail(TCVN_ERR_UNSUPPORTED, "stem: map too large to bin");
  const BinLayout L = bin_layout(n_images, Hb, Wb, nnz);
  char* b = static_cast<char*>(bins);
  int32_t* tile_local = reinterpret_cast<int32_t*>(b + L.local);
  int32_t* totals = reinterpret_cast<int32_t*>(b + L.totals);
  int32_t* base = reinterpret_cast<int32_t*>(b + L.base);
  int32_t* tile_start = reinterpret_cast<int32_t*>(b + L.start);
  float4* recs = reinterpret_cast<float4*>(b + L.recs);
  const size_t smem = (size_t)per_image * sizeof(int);
  if (smem > 48 * 1024) {
    TCVN_CUDA(cudaFuncSetAttribute(stem_bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(stem_bin_fill_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(stem_bin_fill_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  stem_bin_count_kernel<<<n_images, 256, smem, stream>>>(coords, image_offsets, H, W, tiles_x, tiles_y, tile_local, totals);
  TCVN_LAUNCH_CHECK();
  stem_scan_kernel<<<1, 1024, 0, stream>>>(totals, n_images, base);
  TCVN_LAUNCH_CHECK();
  if (values_u8)
    stem_bin_fill_kernel<uint8_t><<<n_images, 32, smem, stream>>>(coords, static_cast<const uint8_t*>(values), image_offsets,
                                                                  n_images, cin, divisor, H, W, tiles_x, tiles_y, tile_local,
                                                                  base, tile_start, recs);
  else
    stem_bin_fill_kernel<float><<<n_images, 32, smem, stream>>>(coords, static_cast<const float*>(values), image_offsets,
                                                                n_images, cin, divisor, H, W, tiles_x, tiles_y, tile_local, base,
                                                                tile_start, recs);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// images [image0, image0 + n) of a binned call -> channels [0, 64) of block 0's ringed buffer (image 0 of blk = image0)
int launch_stem_coo_binned(int image0, int n, int n_images_binned, long long nnz_binned, int cin, int H, int W, const float* w0,
                           const float* s_scale, const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb,
                           int Wb, bool f32, const void* bins, cudaStream_t stream) {
  if (c0 != kC0) return fail(TCVN_ERR_UNSUPPORTED, "stem: init_features %d (kernel is specialised for 64)", c0);
  if (n == 0) return TCVN_OK;
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  if ((Hs - 3) / 2 + 1 != Hb || (Ws - 3) / 2 + 1 != Wb) return fail(TCVN_ERR_ARG, "stem: geometry mismatch");
  if (cin > 3) return fail(TCVN_ERR_UNSUPPORTED, "stem: %d input channels (kernel handles up to 3)", cin);
  const BinLayout L = bin_layout(n_images_binned, Hb, Wb, nnz_binned);
  const char* b = static_cast<const char*>(bins);
  const int32_t* tile_start = reinterpret_cast<const int32_t*>(b + L.start);
  const float4* recs = reinterpret_cast<const float4*>(b + L.recs);
  const long long tiles = (long long)n * tiles_of(Hb) * tiles_of(Wb);
  const size_t smem_max = (size_t)(kTaps + kWarps * kWPos) * kC0 * sizeof(float);
  const int sms = sm_count();   // honours tcvn_set_sm_limit
  static const int warps = getenv("TCVN_STEM_WARPS") ? atoi(getenv("TCVN_STEM_WARPS")) : kWarps;   // experiments: 1..9
  const long long want = (tiles + warps - 1) / warps;
  const int grid = (int)(want < sms ? want : sms);
  const size_t smem = (size_t)(kTaps + warps * kWPos) * kC0 * sizeof(float);
#define TCVN_STEM_WARP(TO)                                                                                              \
  do {                                                                                                                  \
    bool& done = device_flag(f32 ? 6 : 7);                                                                              \
    if (!done) {                                                                                                        \
      TCVN_CUDA(cudaFuncSetAttribute(stem_warp_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max)); \
      done = true;                                                                                                      \
    }                                                                                                                   \
    stem_warp_kernel<TO><<<grid, warps * 32, smem, stream>>>(tile_start, recs, image0, n, cin, w0, s_scale, s_shift,   \
                                                               s_alpha, static_cast<TO*>(blk), ldo, Hb, Wb);            \
  } while (0)
  if (f32) TCVN_STEM_WARP(float);
  else TCVN_STEM_WARP(__nv_bfloat16);
#undef TCVN_STEM_WARP
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn
