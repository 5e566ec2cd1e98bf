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
//     GroupNorm's statistics from the values it stores: (sum, sum^2) per tile in a fixed slot (added in a fixed order by
//     gn_tiles_finalize_kernel: bit-reproducible, no atomics).
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
constexpr int kThreads = 576;                         // warp 0 TMA, warp 1 MMA, warps 2-9 transform, warps 10-17 two epilogue groups
constexpr int kXform = 256;
constexpr uint32_t kDescHiPatch = ((uint32_t)(kPX * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = 1280 B | version | SWIZZLE_128B

struct Conv2dParams {
  int n_img, H, W, Hp, Wp;
  int tiles_x, tiles_y, num_tiles;
  const float* bias;                 // [64]
  const float2* in_stat;             // [n_img] (mean, rstd) of the input's GroupNorm
  const float* gamma; const float* beta;   // [64]
  const bf16* residual;              // optional [n_img * Hp * Wp][64]
  bf16* out;                         // [n_img * Hp * Wp][64]
  double* stat_parts;                // optional [num_tiles][2]
};

__device__ __forceinline__ float bflo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bfhi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float silu(float v) { return __fdividef(v, 1.f + __expf(-v)); }

__global__ void __launch_bounds__(kThreads, 1) umma_conv2d_c64_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                      const Conv2dParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                   // [9][64 n x 128 B]
  uint8_t* sA = smem + kWBytes;                         // [stages][180 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kStages * kStageStride);
  uint64_t* full = bars;                     // TMA -> transform
  uint64_t* ready = bars + kStages;          // transform -> MMA
  uint64_t* empty = bars + 2 * kStages;      // MMA -> TMA
  uint64_t* tfull = bars + 3 * kStages;      // MMA -> epilogue [2]
  uint64_t* tempty = tfull + 2;              // epilogue -> MMA [2]
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 2);   // [2 groups][4 warps][2]
  float* s_bias = s_stat + 16;                                // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&ready[s], kXform);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::mbar_init(wfull, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
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
        const int img = tile / per_img, rem = tile - img * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
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
  } else if (warp < 10) {
    // GroupNorm + SiLU in place.  Thread t owns channel group cg = t & 7 (its gamma / beta stay in registers) on patch rows
    // (t >> 3) + 32 k; under the 128B swizzle those channels sit at 16-byte position cg ^ (row & 7): a warp touches four whole
    // rows per step, conflict-free.
    const int t = threadIdx.x - 64;
    const int cg = t & 7, r0 = t >> 3;
    float ga[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ga[i] = __ldg(p.gamma + cg * 8 + i); be[i] = __ldg(p.beta + cg * 8 + i); }
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const float2 st = __ldg(p.in_stat + img);
      float sc[8], sh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { sc[i] = st.y * ga[i]; sh[i] = fmaf(-st.x, sc[i], be[i]); }
      if (lane == 0) ptx::mbar_wait(&full[stage], phase);
      __syncwarp();
      uint8_t* base = sA + stage * kStageStride;
      for (int r = r0; r < kPatchRows; r += 32) {
        const int py = r / kPX, px = r - py * kPX;
        const int yy = ty * kTY - 1 + py, xx = tx * kTX - 1 + px;       // ringed coordinates of this patch pixel
        uint4* q = reinterpret_cast<uint4*>(base + r * 128 + ((cg ^ (r & 7)) << 4));
        if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) {
          const uint4 v = *q;
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = pack2(silu(fmaf(bflo(w[i]), sc[2 * i], sh[2 * i])), silu(fmaf(bfhi(w[i]), sc[2 * i + 1], sh[2 * i + 1])));
          *q = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
          *q = make_uint4(0u, 0u, 0u, 0u);      // ring / outside the image: the convolution's zero padding
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&ready[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int grp = (warp - 10) >> 2;   // epilogue group == TMEM accumulator it drains
    const int g = warp & 3;             // TMEM lane quarter this warp may read
    const int row = g * 32 + lane;      // pixel of the tile: (row >> 3, row & 7)
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int img = tile / per_img, rem = tile - img * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int yy = ty * kTY + (row >> 3), xx = tx * kTX + (row & 7);
      const bool inb = yy < p.Hp && xx < p.Wp;
      const bool interior = yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
      const size_t prow = ((size_t)img * p.Hp + yy) * p.Wp + xx;
      uint4 res[8];
      if (p.residual != nullptr && interior) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + prow * kC);
#pragma unroll
        for (int j = 0; j < 8; ++j) res[j] = __ldg(rp + j);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) res[j] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (lane == 0) ptx::mbar_wait(&tfull[grp], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + grp * kC, r0);
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + grp * kC + 32, r1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[grp]);
      float s1 = 0.f, s2 = 0.f;
      uint4 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* r = j < 4 ? r0 : r1;
        const int b = (j & 3) * 8;
        const uint32_t rw[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = __uint_as_float(r[b + 2 * i]) + s_bias[j * 8 + 2 * i] + bflo(rw[i]);
          const float hi = __uint_as_float(r[b + 2 * i + 1]) + s_bias[j * 8 + 2 * i + 1] + bfhi(rw[i]);
          w[i] = interior ? pack2(lo, hi) : 0u;
          const float a = bflo(w[i]), c = bfhi(w[i]);     // statistics of exactly the stored bf16 values
          s1 += a + c;
          s2 = fmaf(a, a, fmaf(c, c, s2));
        }
        o[j] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (inb) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + prow * kC);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = o[j];
      }
      if (p.stat_parts != nullptr) {   // warp-uniform
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, d);
          s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        ptx::named_bar_sync(1 + grp, 128);          // the previous tile's slot of this group has been consumed
        if (lane == 0) { s_stat[(grp * 4 + g) * 2] = s1; s_stat[(grp * 4 + g) * 2 + 1] = s2; }
        ptx::named_bar_sync(1 + grp, 128);
        if (g == 0 && lane == 0) {
          double a = 0.0, b = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { a += (double)s_stat[(grp * 4 + k) * 2]; b += (double)s_stat[(grp * 4 + k) * 2 + 1]; }
          p.stat_parts[(size_t)tile * 2] = a;
          p.stat_parts[(size_t)tile * 2 + 1] = b;
        }
      }
      acc_phase ^= 1;
    }
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

int make_map_4d(const void* base, int C, int Wp, int Hp, int n, CUtensorMap* out) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return fail(TCVN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fp);
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)n};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)kPX, (cuuint32_t)kPY, 1};
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
  p.bias = bias; p.in_stat = static_cast<const float2*>(in_stat); p.gamma = gamma; p.beta = beta;
  p.residual = static_cast<const bf16*>(residual_bf16); p.out = static_cast<bf16*>(out_bf16);
  p.stat_parts = static_cast<double*>(stat_parts);
  CUtensorMap tmX, tmW;
  TCVN_TRY(make_map_4d(x_bf16, kC, p.Wp, p.Hp, n, &tmX));
  TCVN_TRY(make_map(w_bf16, 9 * kC, kC, kC, 64, kC, &tmW));
  const size_t smem = 1024 + kWBytes + (size_t)kStages * kStageStride + (3 * kStages + 5) * 8 + 16 + (16 + 64) * 4;
  bool& attr_done = device_flag(5);
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_conv2d_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  umma_conv2d_c64_kernel<<<grid, kThreads, smem, stream>>>(tmX, tmW, p);
  TCVN_LAUNCH_CHECK();
  if (out_stat) {
    gn_tiles_finalize_kernel<<<n, 256, 0, stream>>>(p.stat_parts, p.tiles_x * p.tiles_y, (double)H * W * kC, eps,
                                                    static_cast<float2*>(out_stat));
    TCVN_LAUNCH_CHECK();
  }
  return TCVN_OK;
}
