// 3x3 convolution, 64 -> 64 channels, on tcgen05 with 2-D spatial tiles (sm_100a) - the full-resolution layers of the --sdxl
// pixel-map CNN (BASELINE configs[3]; transformercvn/network/layers/sdxl_net.py:27-34 = diffusers' ResnetBlock2D: 72 % of
// that network's 57 GFLOP per image are 64-channel 3x3 convolutions at 400x280 / 200x140).
//
//   out = conv3x3( silu( GroupNorm_1group(x) ) ) + bias [+ residual]        over ringed channels-last bf16 maps
//
// The shifted-GEMM form of umma.cu (launch_gemm_shifted) reloads the A tile from L2 for each of the nine taps: with a
// 64-channel input that is 33 FLOP per byte of L2 traffic and the kernel ran at 14-17 % of the bf16 peak.  Here a CTA
// loads ONE haloed patch per tile - 18 x 10 pixels x 64 channels (23 KB) through a 4-D tensor map (channels, x, y, image:
// out-of-image pixels are zero-filled by TMA) - and all nine taps read it from shared memory:
//   * the patch is stored as 180 rows of 128 bytes (128B-swizzled, as TMA writes it); an output tile is 16 x 8 pixels, so
//     the eight pixels of one output row are one 8-row core-matrix group, and consecutive groups are TEN rows apart:
//     the UMMA descriptor of tap (dy, dx) starts at row dy * 10 + dx with stride-byte-offset 1280 (the 128B swizzle is a
//     function of the absolute shared-memory address, so any row is a legal start - measured in round 1 on the 1-D tiles);
//   * the 9 x [64 x 64] filter taps (72 KB) stay resident in shared memory; 36 MMAs (M128 N64 K16) per tile;
//   * GroupNorm + SiLU are applied by eight transform warps IN PLACE in the swizzled patch (the patch is loaded once, not
//     nine times, so the transform costs 23 KB per tile): the activated map is never materialised; pixels outside the
//     image (the conv's zero padding) are forced to zero AFTER the activation;
//   * the epilogue adds bias and the ResNet block's residual input, writes zeros on the ring, and takes the NEXT
//     GroupNorm's statistics from the fp32 values it is about to round: (sum, sum^2) per tile in a fixed slot (added in a
//     fixed order by gn_tiles_finalize_kernel: bit-reproducible, no atomics).
// The SIMT side is what bounds the kernel (ncu, profiles/r2_sdxl_conv2d_c64*.txt): the transform and the epilogue work in
// packed fp32 pairs (FFMA2 / FADD2), SiLU costs one MUFU (tanh.approx) instead of two, and everything that does not depend
// on the tile (patch-row coordinates, swizzled slots) is hoisted out of the tile loop.
// A ResNet block of the 64-channel stages is therefore two launches of this kernel and two tiny finalize launches.
#include "ptx.cuh"
#include "umma.h"

namespace tcvn {

using bf16 = __nv_bfloat16;

namespace {

constexpr int kTY = 16, kTX = 8;                      // output tile (pixels)
constexpr int kPY = kTY + 2, kPX = kTX + 2;           // haloed patch
constexpr int kPatchRows = kPY * kPX;                 // 180 rows of 128 bytes
constexpr int kPatchBytes = kPatchRows * 128;         // 23040
constexpr int kStageStride = 23 * 1024;               // patch stages start on 1024-byte boundaries (swizzle atoms)
constexpr int kStages = 4;
constexpr int kC = 64;                                // channels in = channels out
constexpr int kWTap = kC * 128;                       // one filter tap: [64 n][64 k] bf16 = 8 KB
constexpr int kWBytes = 9 * kWTap;                    // 72 KB
constexpr int kThreads = 608;                         // warp 0 TMA, warp 1 MMA, warps 2-9 transform, warps 10-17 two epilogue groups, warp 18 residual TMA
constexpr int kOutBytes = kTY * kTX * 128;            // one output / residual tile: 128 pixels x 128 bytes
constexpr int kXform = 256;
constexpr uint32_t kDescHiPatch = ((uint32_t)(kPX * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = 1280 B | version | SWIZZLE_128B

struct Conv2dParams {
  int n_img, H, W, Hp, Wp;
  int tiles_x, tiles_y, num_tiles;
  unsigned long long magic_img, magic_row;   // ceil(2^40 / tiles per image), ceil(2^40 / tiles_x): q = n * magic >> 40
  const float* bias;                 // [64]
  const float2* in_stat;             // [n_img] (mean, rstd) of the input's GroupNorm
  const float* gamma; const float* beta;   // [64]
  int has_residual;                  // the residual map is read through tmR
  double* stat_parts;                // optional [num_tiles][2]
};

__device__ __forceinline__ float bflo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bfhi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// tile -> (image, tile row, tile column) without integer divisions (exact for n * d < 2^40, checked by the launcher)
__device__ __forceinline__ void tile_coords(const Conv2dParams& p, int tile, int per_img, int& img, int& ty, int& tx) {
  img = (int)(((unsigned long long)(unsigned)tile * p.magic_img) >> 40);
  const int rem = tile - img * per_img;
  ty = (int)(((unsigned long long)(unsigned)rem * p.magic_row) >> 40);
  tx = rem - ty * p.tiles_x;
}

__device__ __forceinline__ float tanh_approx(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v));
  return t;
}
// silu(y) = y * sigmoid(y) = h + h * tanh(h), h = y / 2: one MUFU per value instead of two (ex2 + rcp); tanh.approx is good
// to 2^-11 relative, i.e. the result is within 2.4e-4 |y| - a sixteenth of the bf16 rounding that follows
__device__ __forceinline__ float2 silu_half2(float2 h) {
  return __ffma2_rn(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
}

__global__ void __launch_bounds__(kThreads, 1) umma_conv2d_c64_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                      const __grid_constant__ CUtensorMap tmR,
                                                                      const __grid_constant__ CUtensorMap tmO,
                                                                      const Conv2dParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                   // [9][64 n x 128 B]
  uint8_t* sA = smem + kWBytes;                         // [stages][180 rows x 128 B]
  uint8_t* sO = sA + kStages * kStageStride;            // [2 epilogue groups][128 pixels x 128 B]: residual in, result out
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * kOutBytes);
  uint64_t* full = bars;                     // TMA -> transform
  uint64_t* ready = bars + kStages;          // transform -> MMA
  uint64_t* empty = bars + 2 * kStages;      // MMA -> TMA
  uint64_t* tfull = bars + 3 * kStages;      // MMA -> epilogue [2]
  uint64_t* tempty = tfull + 2;              // epilogue -> MMA [2]
  uint64_t* wfull = tempty + 2;
  uint64_t* rfull = wfull + 1;               // residual TMA -> epilogue [2]
  uint64_t* rempty = rfull + 2;              // epilogue (tile stored) -> residual TMA [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + 2);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 2);   // [2 tile parities][2 groups][4 warps][2]
  float* s_bias = s_stat + 32;                                // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&ready[s], kXform);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128);
      ptx::mbar_init(&rfull[a], 1); ptx::mbar_init(&rempty[a], 1);
    }
    ptx::mbar_init(wfull, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmO);
  }
  if (threadIdx.x < kC) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wfull, kWBytes);
      for (int t = 0; t < 9; ++t) ptx::tma_load_2d(sW + t * kWTap, &tmW, wfull, 0, t * kC);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int img, ty, tx;
        tile_coords(p, tile, per_img, img, ty, tx);
        ptx::mbar_wait_long(&empty[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&full[stage], kPatchBytes);
        // patch origin = output tile origin - 1 in ringed coordinates (may be -1: zero-filled)
        ptx::tma_load_4d(sA + stage * kStageStride, &tmX, &full[stage], 0, tx * kTX - 1, ty * kTY - 1, img);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, kC);
      ptx::mbar_wait(wfull, 0);
      const uint32_t w_lo = ptx::umma_desc_lo(ptx::smem_u32(sW));
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::mbar_wait(&ready[stage], phase);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kC;
        const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(sA + stage * kStageStride));
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t row_lo = a_lo + (uint32_t)((t / 3) * kPX + (t % 3)) * 8u;   // +8 (x16 B) per patch row
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(d_tmem, ptx::umma_desc_join(kDescHiPatch, row_lo + 2u * k),
                           ptx::umma_desc_join(ptx::kUmmaDescHiSw128, w_lo + (uint32_t)t * (kWTap >> 4) + 2u * k), idesc, (t | k) != 0);
        }
        ptx::umma_commit(&empty[stage]);
        ptx::umma_commit(&tfull[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else if (warp == 18) {
    // residual tiles -> the epilogue groups' staging buffers (the buffer is free once the previous tile's TMA store has read it)
    if (lane == 0) {
      uint32_t ph[2] = {0u, 0u};
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int grp = it & 1;
        int img, ty, tx;
        tile_coords(p, tile, per_img, img, ty, tx);
        ptx::mbar_wait_long(&rempty[grp], ph[grp] ^ 1);
        if (p.has_residual) {
          ptx::mbar_arrive_expect_tx(&rfull[grp], kOutBytes);
          ptx::tma_load_4d(sO + grp * kOutBytes, &tmR, &rfull[grp], 0, tx * kTX, ty * kTY, img);
        } else {
          ptx::mbar_arrive(&rfull[grp]);
        }
        ph[grp] ^= 1;
      }
    }
  } else if (warp < 10) {
    // GroupNorm + SiLU in place.  Thread t owns channel group cg = t & 7 (its gamma / beta stay in registers) on patch rows
    // (t >> 3) + 32 k; under the 128B swizzle those channels sit at 16-byte position cg ^ (row & 7): a warp touches four whole
    // rows per step, conflict-free.
    const int t = threadIdx.x - 64;
    const int cg = t & 7, r0 = t >> 3;
    float2 ga[4], be[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ga[i] = make_float2(__ldg(p.gamma + cg * 8 + 2 * i), __ldg(p.gamma + cg * 8 + 2 * i + 1));
      be[i] = make_float2(__ldg(p.beta + cg * 8 + 2 * i), __ldg(p.beta + cg * 8 + 2 * i + 1));
    }
    // this thread's patch rows r0 + 32 k never change: (py, px) and the swizzled 16-byte slot of each are loop constants
    constexpr int kIter = (kPatchRows + 31) / 32;
    int pyx[kIter], slot[kIter];
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
      const int r = r0 + 32 * k;
      pyx[k] = r < kPatchRows ? ((r / kPX) << 8) | (r % kPX) : -1;
      slot[k] = r * 128 + ((cg ^ (r & 7)) << 4);
    }
    int stage = 0; uint32_t phase = 0;
    int tile = blockIdx.x;
    float2 st = make_float2(0.f, 0.f);
    if (tile < p.num_tiles) st = __ldg(p.in_stat + (int)(((unsigned long long)(unsigned)tile * p.magic_img) >> 40));
    for (; tile < p.num_tiles; tile += gridDim.x) {
      int img, ty, tx;
      tile_coords(p, tile, per_img, img, ty, tx);
      // (mean, rstd) of the NEXT tile's image is fetched under this tile's work
      const int nxt = tile + gridDim.x;
      float2 st_next = make_float2(0.f, 0.f);
      if (nxt < p.num_tiles) st_next = __ldg(p.in_stat + (int)(((unsigned long long)(unsigned)nxt * p.magic_img) >> 40));
      float2 hsc[4], hsh[4];   // y / 2 = v * hsc + hsh
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hsc[i] = make_float2(0.5f * st.y * ga[i].x, 0.5f * st.y * ga[i].y);
        hsh[i] = make_float2(fmaf(-st.x, hsc[i].x, 0.5f * be[i].x), fmaf(-st.x, hsc[i].y, 0.5f * be[i].y));
      }
      const int y0 = ty * kTY - 1, x0 = tx * kTX - 1;          // ringed coordinates of the patch origin
      if (lane == 0) ptx::mbar_wait(&full[stage], phase);
      __syncwarp();
      const uint32_t base = ptx::smem_u32(sA + stage * kStageStride);
      auto activate = [&](uint4 v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 y = silu_half2(__ffma2_rn(make_float2(bflo(w[i]), bfhi(w[i])), hsc[i], hsh[i]));
          u[i] = pack2(y.x, y.y);
        }
        return make_uint4(u[0], u[1], u[2], u[3]);
      };
      if (y0 >= 1 && y0 + kPY - 1 <= p.H && x0 >= 1 && x0 + kPX - 1 <= p.W) {
        // the whole patch lies inside the image (most tiles): no per-pixel tests
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
          if (pyx[k] < 0) continue;
          ptx::sts128(base + slot[k], activate(ptx::lds128(base + slot[k])));
        }
      } else {
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
          if (pyx[k] < 0) continue;
          const int yy = y0 + (pyx[k] >> 8), xx = x0 + (pyx[k] & 255);
          uint4 o = make_uint4(0u, 0u, 0u, 0u);       // ring / outside the image: the convolution's zero padding
          if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) o = activate(ptx::lds128(base + slot[k]));
          ptx::sts128(base + slot[k], o);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&ready[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
      st = st_next;
    }
  } else {
    const int grp = (warp - 10) >> 2;   // epilogue group == TMEM accumulator it drains
    const int g = warp & 3;             // TMEM lane quarter this warp may read
    const int row = g * 32 + lane;      // pixel of the tile: (row >> 3, row & 7)
    // this pixel's 128 bytes of the staging tile (128B-swizzled like every TMA tile): chunk j sits at j ^ (row & 7)
    const uint32_t srow = ptx::smem_u32(sO + grp * kOutBytes + row * 128);
    const int sw = row & 7;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      int img, ty, tx;
      tile_coords(p, tile, per_img, img, ty, tx);
      const int yy = ty * kTY + (row >> 3), xx = tx * kTX + (row & 7);
      const bool interior = yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
      if (lane == 0) ptx::mbar_wait_long(&rfull[grp], acc_phase);   // staging tile free (and the residual landed in it)
      __syncwarp();
      if (lane == 0) ptx::mbar_wait_long(&tfull[grp], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      // bias + residual in packed fp32 pairs; the NEXT GroupNorm's statistics are taken from the fp32 values (the bf16
      // rounding of the stored copy is zero-mean noise three orders below the variance).  Two halves of 32 channels keep
      // the register footprint under the 96 this 19-warp CTA allows.
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + grp * kC + half * 32, r);
        ptx::tmem_ld_wait();
        if (half == 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty[grp]);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = half * 4 + jj;
          const uint32_t slot = srow + ((j ^ sw) << 4);
          const uint4 res = p.has_residual ? ptx::lds128(slot) : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + j * 8), b1 = *reinterpret_cast<const float4*>(s_bias + j * 8 + 4);
          const float2 bias2[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 v = make_float2(__uint_as_float(r[jj * 8 + 2 * i]), __uint_as_float(r[jj * 8 + 2 * i + 1]));
            v = __fadd2_rn(v, bias2[i]);
            v = __fadd2_rn(v, make_float2(bflo(rw[i]), bfhi(rw[i])));
            w[i] = pack2(v.x, v.y);
            s1 = __fadd2_rn(s1, v);
            s2 = __ffma2_rn(v, v, s2);
          }
          // ring pixels store zeros; pixels beyond the map are clipped by the TMA store
          ptx::sts128(slot, interior ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0u, 0u, 0u, 0u));
        }
      }
      if (!interior) { s1 = make_float2(0.f, 0.f); s2 = make_float2(0.f, 0.f); }   // ... and count for nothing
      float* stat_slot = s_stat + (((it >> 1) & 1) * 2 + grp) * 8;   // double-buffered by this group's tile parity
      if (p.stat_parts != nullptr) {   // warp-uniform
        float t1 = s1.x + s1.y, t2 = s2.x + s2.y;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          t1 += __shfl_xor_sync(0xffffffffu, t1, d);
          t2 += __shfl_xor_sync(0xffffffffu, t2, d);
        }
        if (lane == 0) { stat_slot[g * 2] = t1; stat_slot[g * 2 + 1] = t2; }
      }
      ptx::fence_proxy_async_smem();              // the tile rows written above -> visible to the TMA store
      ptx::named_bar_sync(1 + grp, 128);
      if (g == 0 && lane == 0) {
        ptx::tma_store_4d(&tmO, sO + grp * kOutBytes, 0, tx * kTX, ty * kTY, img);
        ptx::tma_store_commit();
        if (p.stat_parts != nullptr) {
          double a = 0.0, b = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { a += (double)stat_slot[k * 2]; b += (double)stat_slot[k * 2 + 1]; }
          p.stat_parts[(size_t)tile * 2] = a;
          p.stat_parts[(size_t)tile * 2 + 1] = b;
        }
        ptx::tma_store_wait_read();               // the store has read the tile out of shared memory
        ptx::mbar_arrive(&rempty[grp]);           // -> the next residual tile may land
      }
      acc_phase ^= 1;
    }
    if (g == 0 && lane == 0) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 128);
}

// stat[img] = (mean, rstd) from the per-tile partial sums of the producing convolution, tiles added in a fixed order
__global__ void __launch_bounds__(256) gn_tiles_finalize_kernel(const double* __restrict__ parts, int tiles_per_img, double count, float eps,
                                                                float2* __restrict__ stat) {
  __shared__ double sh[2][256];
  const int img = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int t = threadIdx.x; t < tiles_per_img; t += 256) {
    a += parts[((size_t)img * tiles_per_img + t) * 2];
    b += parts[((size_t)img * tiles_per_img + t) * 2 + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh[0][0] / count;
    double var = sh[1][0] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[img] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_4d(const void* base, int C, int Wp, int Hp, int n, int box_x, int box_y, CUtensorMap* out) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return fail(TCVN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fp);
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)n};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TCVN_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%d) C=%d Wp=%d Hp=%d n=%d", (int)r, C, Wp, Hp, n);
  return TCVN_OK;
}

}  // namespace
}  // namespace tcvn

using namespace tcvn;

extern "C" size_t tcvn_sdxl16_conv2d_stat_bytes(int n, int H, int W) {
  return (size_t)n * ceil_div(H + 2, kTY) * ceil_div(W + 2, kTX) * 2 * sizeof(double);
}

// out = conv3x3(silu(GroupNorm(x))) + bias (+ residual), 64 -> 64 channels, ringed bf16 maps [n][(H+2)(W+2)][64].
//   in_stat   [n] float2 (mean, rstd) of x's GroupNorm(1 group);  gamma / beta [64]
//   w_bf16    [9 * 64][64]: row t * 64 + n, column k  =  weight[n][k][t / 3][t % 3]
//   stat_parts (nullable) tcvn_sdxl16_conv2d_stat_bytes(): per-tile (sum, sum^2) of out; out_stat (nullable) [n] float2 = the
//   (mean, rstd) of out's GroupNorm, finalised from them (eps)
extern "C" int tcvn_sdxl16_conv2d_c64(const void* x_bf16, int n, int H, int W, const void* in_stat, const float* gamma, const float* beta,
                                      const void* w_bf16, const float* bias, const void* residual_bf16, void* out_bf16,
                                      void* stat_parts, void* out_stat, float eps, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(x_bf16 && in_stat && gamma && beta && w_bf16 && bias && out_bf16 && n >= 0 && H >= 1 && W >= 1,
                 "sdxl16_conv2d_c64: bad arguments");
  TCVN_CHECK_ARG(out_stat == nullptr || stat_parts != nullptr, "sdxl16_conv2d_c64: out_stat needs stat_parts");
  if (n == 0) return TCVN_OK;
  Conv2dParams p{};
  p.n_img = n; p.H = H; p.W = W; p.Hp = H + 2; p.Wp = W + 2;
  p.tiles_x = ceil_div(p.Wp, kTX); p.tiles_y = ceil_div(p.Hp, kTY);
  const long long tiles = (long long)n * p.tiles_x * p.tiles_y;
  if (tiles >= (1ll << 31)) return fail(TCVN_ERR_UNSUPPORTED, "sdxl16_conv2d_c64: too many tiles in one launch");
  p.num_tiles = (int)tiles;
  const unsigned long long per_img = (unsigned long long)p.tiles_x * p.tiles_y;
  if ((unsigned long long)tiles * per_img >= (1ull << 40)) return fail(TCVN_ERR_UNSUPPORTED, "sdxl16_conv2d_c64: too many tiles in one launch");
  p.magic_img = ((1ull << 40) + per_img - 1) / per_img;
  p.magic_row = ((1ull << 40) + p.tiles_x - 1) / p.tiles_x;
  p.bias = bias; p.in_stat = static_cast<const float2*>(in_stat); p.gamma = gamma; p.beta = beta;
  p.has_residual = residual_bf16 != nullptr;
  p.stat_parts = static_cast<double*>(stat_parts);
  CUtensorMap tmX, tmW, tmR, tmO;
  TCVN_TRY(make_map_4d(x_bf16, kC, p.Wp, p.Hp, n, kPX, kPY, &tmX));
  TCVN_TRY(make_map_4d(residual_bf16 ? residual_bf16 : out_bf16, kC, p.Wp, p.Hp, n, kTX, kTY, &tmR));
  TCVN_TRY(make_map_4d(out_bf16, kC, p.Wp, p.Hp, n, kTX, kTY, &tmO));
  TCVN_TRY(make_map(w_bf16, 9 * kC, kC, kC, 64, kC, &tmW));
  const size_t smem = 1024 + kWBytes + (size_t)kStages * kStageStride + 2 * kOutBytes + (3 * kStages + 9) * 8 + 16 + (32 + 64) * 4;
  bool& attr_done = device_flag(5);
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_conv2d_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  umma_conv2d_c64_kernel<<<grid, kThreads, smem, stream>>>(tmX, tmW, tmR, tmO, p);
  TCVN_LAUNCH_CHECK();
  if (out_stat) {
    gn_tiles_finalize_kernel<<<n, 256, 0, stream>>>(p.stat_parts, p.tiles_x * p.tiles_y, (double)H * W * kC, eps,
                                                    static_cast<float2*>(out_stat));
    TCVN_LAUNCH_CHECK();
  }
  return TCVN_OK;
}
