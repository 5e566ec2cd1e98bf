// tcgen05 kernels of the TRAINING path (bf16 operands, fp32 accumulation in TMEM), sm_100a.
//
//   umma_wgrad_kernel      weight gradients: dW[item][k][n] = sum_m act(A[m + shift_item, col_item + k]) * G[m, n]
//       The reduction runs over PIXEL ROWS, which are the slow dimension of both row-major operands, so both
//       MMA operands are MN-major: TMA drops a [64 rows x 64 channels] box (128-byte rows, 128B swizzle) and the
//       instruction descriptor's a_major/b_major bits make the tensor core read it transposed - no transposed copy
//       of an activation ever exists.  M = 128 channels of A (two 64-channel boxes, LBO apart), N = 128 channels
//       of G, K = 16 rows per instruction.  Up to four "items" (column blocks of A, or row-shifted views of A)
//       accumulate into four 128-column TMEM accumulators that live for the whole kernel; each persistent CTA owns a
//       strided set of 64-row tiles and adds its partial sums to the fp32 result with red.global at the end.
//         conv1:  items = 128-channel column blocks of the concat buffer, act = BN1 + PReLU1 applied in place in
//                 the swizzled SMEM tile (same trick as the forward conv1 kernel), G = d(mid)
//         conv2:  items = the three vertical taps (row shifts -Wp, 0, +Wp) of the activated bottleneck map,
//                 G = the gradient of the 32 output channels in its three horizontal shifts (G2x, 96 + 32 zero
//                 columns) so that one N=128 MMA covers a whole filter row
//   umma_conv2_dgrad_kernel   d(act mid)[p][c] = sum_dy sum_k G2x[p + (1-dy) Wp][k] * Wd[dy][c][k]
//       3x3 input gradient as a 3-tap shifted GEMM over one haloed tile (K-major, like the forward conv2 kernel), the
//       horizontal taps being the K blocks of G2x.  N = 128.
#include "ptx.cuh"
#include "umma.h"

namespace tcvn {

using bf16 = __nv_bfloat16;

namespace {

constexpr int kWgRows = 64;                 // pixel rows per stage
constexpr int kWgBox = kWgRows * 128;       // 8 KB: one [64 rows x 64 channels] box
constexpr int kWgThreads = 448;             // warp 0 TMA, warp 1 MMA, warps 2-9 transform, warps 10-13 epilogue
constexpr int kWgXform = 256;

struct WgradParams {
  long long rows;
  int n_items;            // 1..4
  int item_col[4];        // first A column of the item (elements)
  int item_shift[4];      // row shift of the item's A view
  int item_valid[4];      // A columns of the item that are real (<= 128): rows of the result to write
  int a_cols;             // columns of A covered by the transform constants (channels >= a_cols are forced to 0)
  const float *a_scale, *a_shift, *a_alpha;
  int g_col0;
  float* dw;              // [gridDim.x][n_items][128][128] fp32: one partial sum per CTA (plain stores)
  int num_tiles, stages;
};

__device__ __forceinline__ float bfl(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bfh(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

constexpr uint32_t kIdescMN = (1u << 15) | (1u << 16);   // a_major = b_major = MN

template <bool TRANSFORM>
__global__ void __launch_bounds__(kWgThreads, 1) umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmG,
                                                                   const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (2 * p.n_items + 2) * kWgBox;   // A boxes of every item, then the two G boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* ready = bars + 8;
  uint64_t* empty = bars + 16;
  uint64_t* done = bars + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&ready[s], kWgXform);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmG);
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row0 = tile * kWgRows;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&full[stage], stage_bytes);
        uint8_t* s = smem + stage * stage_bytes;
        for (int it = 0; it < p.n_items; ++it)
          for (int h = 0; h < 2; ++h)
            ptx::tma_load_2d(s + (2 * it + h) * kWgBox, &tmA, &full[stage], p.item_col[it] + h * 64, row0 + p.item_shift[it]);
        for (int h = 0; h < 2; ++h)
          ptx::tma_load_2d(s + (2 * p.n_items + h) * kWgBox, &tmG, &full[stage], p.g_col0 + h * 64, row0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 128) | kIdescMN;
      // MN-major, 128B swizzle: 64 channels contiguous (128 B), the next 64 channels LBO = one box further, 8-row
      // groups SBO = 1024 B apart; one instruction consumes 16 rows = 2048 B
      const uint32_t lbo = ((uint32_t)kWgBox >> 4) << 16;
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(TRANSFORM ? &ready[stage] : &full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t s_addr = ptx::smem_u32(smem + stage * stage_bytes);
        const uint32_t g_lo = (((s_addr + 2 * p.n_items * kWgBox) & 0x3FFFFu) >> 4) | lbo;
#pragma unroll 1
        for (int it = 0; it < p.n_items; ++it) {
          const uint32_t a_lo = (((s_addr + 2 * it * kWgBox) & 0x3FFFFu) >> 4) | lbo;
#pragma unroll
          for (int ks = 0; ks < kWgRows / 16; ++ks)
            ptx::umma_bf16(tmem_base + it * 128, ptx::umma_desc_join(ptx::kUmmaDescHiSw128, a_lo + ks * 128),
                           ptx::umma_desc_join(ptx::kUmmaDescHiSw128, g_lo + ks * 128), idesc, !(first && ks == 0));
        }
        first = false;
        ptx::umma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(done);
    }
  } else if (warp < 10) {
    if (TRANSFORM) {
      // warp w owns channel group cg = w of every box, lane l its rows l and l + 32 (16-byte position cg ^ (l & 7) under
      // the 128B swizzle): conflict-free, and the constant loads below are warp-wide broadcasts (with all 8 groups in one
      // warp they cost more LSU wavefronts than the operand transform itself, cf. umma_gemm_kernel)
      const int t = threadIdx.x - 64;
      const int cg = t >> 5;
      const int rb = lane, pc = cg ^ (lane & 7);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (lane == 0) ptx::mbar_wait(&full[stage], phase);
        __syncwarp();
        uint8_t* s = smem + stage * stage_bytes;
        for (int bx = 0; bx < 2 * p.n_items; ++bx) {
          const int ch = p.item_col[bx >> 1] + (bx & 1) * 64 + cg * 8;
          uint8_t* base = s + bx * kWgBox + rb * 128 + pc * 16;
          if (ch < p.a_cols) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(p.a_scale + ch)), x1 = __ldg(reinterpret_cast<const float4*>(p.a_scale + ch) + 1);
            const float4 y0 = __ldg(reinterpret_cast<const float4*>(p.a_shift + ch)), y1 = __ldg(reinterpret_cast<const float4*>(p.a_shift + ch) + 1);
            const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.a_alpha + ch)), z1 = __ldg(reinterpret_cast<const float4*>(p.a_alpha + ch) + 1);
            const float sc[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float sh[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            const float al[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 v = *reinterpret_cast<const uint4*>(base + i * 32 * 128);
              uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __nv_bfloat162 y = __floats2bfloat162_rn(fmaf(bfl(w[q]), sc[2 * q], sh[2 * q]),
                                                               fmaf(bfh(w[q]), sc[2 * q + 1], sh[2 * q + 1]));
                const __nv_bfloat162 a2 = __floats2bfloat162_rn(al[2 * q], al[2 * q + 1]);
                const __nv_bfloat162 r = __hfma2(a2, __hmin2(y, zero2), __hmax2(y, zero2));
                w[q] = *reinterpret_cast<const uint32_t*>(&r);
              }
              *reinterpret_cast<uint4*>(base + i * 32 * 128) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) *reinterpret_cast<uint4*>(base + i * 32 * 128) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&ready[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // epilogue: after the CTA's last tile, store its partial sums (rows beyond the item's valid channels as zeros);
    // a reduction kernel adds the per-CTA partials - 148 x 64 K atomics per item cost more than the main loop
    const int g = warp & 3;          // TMEM lane group this warp may read (warps 10..13 -> 2,3,0,1)
    const int row = g * 32 + lane;   // channel of A within the item
    if (blockIdx.x < p.num_tiles) {
      if (lane == 0) ptx::mbar_wait(done, 0);
      __syncwarp();
      ptx::tc_fence_after();
      float* part = p.dw + (size_t)blockIdx.x * p.n_items * 128 * 128;
      for (int it = 0; it < p.n_items; ++it) {
        const bool valid = row < p.item_valid[it];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + it * 128 + c * 32, r);
          ptx::tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(part + ((size_t)it * 128 + row) * 128 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = valid ? make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// conv2 input gradient.  One stage = one 64-column half of the haloed G2x tile.
// ------------------------------------------------------------------------------------------------
constexpr int kDgThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 two epilogue groups (one per TMEM accumulator)
constexpr int kDgStages = 3;
constexpr int kDgBoxRows = 32;
constexpr int kDgWSlab = 128 * 128;            // [128 out channels x 64 k] bf16
constexpr int kDgWBytes = 3 * 2 * kDgWSlab;    // 96 KB: [dy][half]

struct DgradParams {
  long long m_total;
  int Hp, Wp, halo_rows, nbox;
  bf16* out; int ldo;
  int num_tiles;
  // BatchNorm2 + PReLU2 backward reductions fused into the epilogue (training walk): with x = the raw conv1 output
  // (bf16 [rows][128]) and fold = [scale | shift | alpha | mean | rstd] x 128 of BN2, every CTA stores
  //   red_parts[cta][3][128] = (sum g, sum g xhat, sum d min(y, 0)),  y = scale x + shift, g = d (y >= 0 ? 1 : alpha),
  // over its rows, d = the bf16 value this kernel stores.  The consumer adds the CTAs in a fixed order (no atomics).
  // Replaces a separate pass over d and x (512 bytes per row).
  const bf16* red_x; const float* red_fold; double* red_parts;
};

__global__ void __launch_bounds__(kDgThreads, 1) umma_conv2_dgrad_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                         const __grid_constant__ CUtensorMap tmW,
                                                                         const DgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int half_bytes = p.halo_rows * 128;
  uint8_t* sW = smem;
  uint8_t* sA = smem + kDgWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kDgStages * half_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDgStages;
  uint64_t* tfull = bars + 2 * kDgStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  __shared__ float4 s_red[128];     // fused BN2 backward reductions: (scale, shift, rstd, -mean * rstd) per column
  __shared__ float s_red_al[128];   // ... and the PReLU slope

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.red_parts != nullptr && threadIdx.x < 128) {
    const int c = threadIdx.x;
    const float sc = p.red_fold[c], sh = p.red_fold[128 + c], mu = p.red_fold[3 * 128 + c], rs = p.red_fold[4 * 128 + c];
    s_red[c] = make_float4(sc, sh, rs, -mu * rs);
    s_red_al[c] = p.red_fold[2 * 128 + c];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDgStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::mbar_init(wfull, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmG);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wfull, kDgWBytes);
      for (int dy = 0; dy < 3; ++dy)
        for (int h = 0; h < 2; ++h)
          ptx::tma_load_2d(sW + (dy * 2 + h) * kDgWSlab, &tmW, wfull, h * 64, dy * 128);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row0 = tile * 128 - p.Wp;   // first row of the haloed tile
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage], half_bytes);
          for (int b = 0; b < p.nbox; ++b)
            ptx::tma_load_2d(sA + stage * half_bytes + b * kDgBoxRows * 128, &tmG, &full[stage], h * 64, row0 + b * kDgBoxRows);
          if (++stage == kDgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
      ptx::mbar_wait(wfull, 0);
      const uint32_t w_lo = ptx::umma_desc_lo(ptx::smem_u32(sW));
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 128;
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(sA + stage * half_bytes));
          const int ksteps = h == 0 ? 4 : 2;   // columns 96..127 of G2x are zero padding
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // out row p reads G2x row p + (1 - dy) * Wp = tile row (p - m0) + (2 - dy) * Wp
            const uint32_t row_lo = a_lo + (uint32_t)((2 - dy) * p.Wp) * 8u;
            for (int kk = 0; kk < ksteps; ++kk)
              ptx::umma_bf16(d_tmem, ptx::umma_desc_join(ptx::kUmmaDescHiSw128, row_lo + kk * 2u),
                             ptx::umma_desc_join(ptx::kUmmaDescHiSw128, w_lo + (dy * 2 + h) * (kDgWSlab >> 4) + kk * 2u), idesc,
                             (h | dy | kk) != 0);
          }
          ptx::umma_commit(&empty[stage]);
          if (++stage == kDgStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull[acc]);
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    // two epilogue groups of four warps, group == the TMEM accumulator it drains (tiles alternate)
    const int grp = (warp - 2) >> 2;
    const int g = warp & 3;
    const int row = g * 32 + lane;
    const int R = p.Hp * p.Wp;
    const int acc = grp; uint32_t acc_phase = 0;
    const bool red = p.red_parts != nullptr;
    float racc[4][3];   // lane j: column c * 32 + j of the three reductions over this warp's rows of every tile
#pragma unroll
    for (int c = 0; c < 4; ++c) { racc[c][0] = 0.f; racc[c][1] = 0.f; racc[c][2] = 0.f; }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const long long m = (long long)tile * 128 + row;
      bool ring = false;
      {
        const int rr = (int)(m % R);
        const int y = rr / p.Wp, x = rr - y * p.Wp;
        ring = y == 0 || y == p.Hp - 1 || x == 0 || x == p.Wp - 1;
      }
      const bool live = m < p.m_total && !ring;
      if (lane == 0) ptx::mbar_wait(&tfull[acc], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(g * 32) << 16) + acc * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_addr + c * 32, r);
        uint4 xr[4];
        if (red && live) {
          const uint4* xp = reinterpret_cast<const uint4*>(p.red_x + m * 128 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) xr[j] = __ldg(xp + j);
        }
        ptx::tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          w[j] = ring ? 0u : *reinterpret_cast<const uint32_t*>(&h2);
        }
        if (m < p.m_total) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + m * p.ldo + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        if (red) {   // warp-uniform
          float q0[32], q1[32], q2[32];
          const uint32_t xw[16] = {xr[0].x, xr[0].y, xr[0].z, xr[0].w, xr[1].x, xr[1].y, xr[1].z, xr[1].w,
                                   xr[2].x, xr[2].y, xr[2].z, xr[2].w, xr[3].x, xr[3].y, xr[3].z, xr[3].w};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 k = s_red[c * 32 + j];              // (scale, shift, rstd, -mean * rstd): warp-wide broadcast
            const float al = s_red_al[c * 32 + j];
            const uint32_t dw = w[j >> 1], xx = xw[j >> 1];
            const float d = live ? ((j & 1) ? bfh(dw) : bfl(dw)) : 0.f;
            const float x = live ? ((j & 1) ? bfh(xx) : bfl(xx)) : 0.f;
            const float y = fmaf(x, k.x, k.y);
            const float gg = y >= 0.f ? d : d * al;
            q0[j] = gg;
            q1[j] = gg * fmaf(x, k.z, k.w);
            q2[j] = d * fminf(y, 0.f);
          }
          racc[c][0] += warp_transpose_sum(q0, lane);
          racc[c][1] += warp_transpose_sum(q1, lane);
          racc[c][2] += warp_transpose_sum(q2, lane);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[acc]);
      acc_phase ^= 1;
    }
    if (red) {
      // the eight epilogue warps' partial sums are added in a fixed order through the (now idle) pipeline stages
      float* sred = reinterpret_cast<float*>(sA);   // [8 warps][3][128]
      const int wi = grp * 4 + g;
      ptx::named_bar_sync(1, 256);                  // every MMA of this CTA has completed: the stages are free
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < 3; ++q) sred[(wi * 3 + q) * 128 + c * 32 + lane] = racc[c][q];
      ptx::named_bar_sync(1, 256);
      for (int o = threadIdx.x - 64; o < 3 * 128; o += 256) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += (double)sred[k * 3 * 128 + o];
        p.red_parts[(size_t)blockIdx.x * 3 * 128 + o] = t;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

// out (+)= sum over the per-CTA partial sums.  A block owns 32 consecutive float4 columns; its 8 warps each sum every
// 8th part (coalesced 512-byte reads, ~18 independent loads per thread), then the 8 slice sums are added in a fixed
// order through shared memory - deterministic, and 8x more blocks / loads in flight than one thread per column
// walking all parts (which left a 16 K-element reduction on 16 blocks: 14-20 us, 8 % of the 16-event training step).
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float4* __restrict__ parts, int n_parts, long long n4,
                                                           float4* __restrict__ out, int accumulate) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
#pragma unroll 4
    for (int q = slice; q < n_parts; q += 8) {
      const float4 v = __ldg(parts + (size_t)q * n4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sh[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && i < n4) {
    float4 t = accumulate ? out[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = sh[k][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    out[i] = t;
  }
}

// Same reduction, fused with the re-layout into the reference's weight-gradient tensor (one launch instead of reduce +
// unpack): element e of the summed [.. ][128][128] result is ADDED to dst at
//   mode 1 (1x1 convolution, items = 128-channel column blocks): e = kp * 128 + n  ->  dst[n][k], k = logical channel of
//          the physical concat channel kp (alignment-padding channels [c0, c0p) are skipped), n < n_log, k < k_log
//   mode 2 (3x3 convolution): e = (dy * 128 + k) * 128 + dx * 32 + n  ->  dst[n][k][dy][dx]  (columns >= 96 are padding)
struct MapArgs { int mode, n_log, k_log, c0, c0p; };
__global__ void __launch_bounds__(256) reduce_parts_map_kernel(const float4* __restrict__ parts, int n_parts, long long n4,
                                                               const MapArgs a, float* __restrict__ dst) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
#pragma unroll 4
    for (int q = slice; q < n_parts; q += 8) {
      const float4 v = __ldg(parts + (size_t)q * n4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sh[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && i < n4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = sh[k][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = i * 4 + j;
      const int col = (int)(e & 127);
      const int row = (int)(e >> 7);
      if (a.mode == 1) {
        int k = -1;
        if (row < a.c0) k = row;
        else if (row >= a.c0p) k = row - (a.c0p - a.c0);
        if (k >= 0 && k < a.k_log && col < a.n_log) dst[(size_t)col * a.k_log + k] += tv[j];
      } else {
        const int dy = row >> 7, k = row & 127, dx = col >> 5, n = col & 31;
        if (dx < 3 && dy < 3) dst[(((size_t)n * 128 + k) * 3 + dy) * 3 + dx] += tv[j];
      }
    }
  }
}

}  // namespace

// out[i] (+)= sum_q parts[q][i], q in a fixed order (n floats per part, n % 4 == 0)
int reduce_parts(const float* parts, int n_parts, long long n, float* out, bool accumulate, cudaStream_t st) {
  if (n % 4) return fail(TCVN_ERR_ARG, "reduce_parts: %lld floats (must be a multiple of 4)", n);
  const long long n4 = n / 4;
  reduce_parts_kernel<<<(unsigned)ceil_div_ll(n4, 32), 256, 0, st>>>(reinterpret_cast<const float4*>(parts), n_parts, n4,
                                                                     reinterpret_cast<float4*>(out), accumulate ? 1 : 0);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

constexpr int kWgMaxCtas = 148;  // per-CTA partial sums: the scratch is sized for this many
size_t umma_wgrad_parts_bytes(int n_items) { return (size_t)kWgMaxCtas * n_items * 128 * 128 * sizeof(float); }

// dw[item][k][n] (fp32 [n_items][128][128]) (+)= sum_m act(A[m + shift_item, col_item + k]) * G[m, g_col0 + n]
// parts: scratch of umma_wgrad_parts_bytes(n_items) for the per-CTA partial sums
int umma_wgrad(const void* A, long long rows, int a_cols, int a_pitch, int n_items, const int* item_col, const int* item_shift,
               const int* item_valid, const float* a_scale, const float* a_shift, const float* a_alpha, int a_fold_cols,
               const void* G, int g_cols, int g_pitch, int g_col0, float* parts, float* dw, bool accumulate, cudaStream_t st,
               int map_mode, int map_n_log, int map_k_log, int map_c0, int map_c0p) {
  if (n_items < 1 || n_items > 4) return fail(TCVN_ERR_ARG, "umma_wgrad: %d items (1..4)", n_items);
  if (rows <= 0) {
    if (!accumulate && map_mode == 0) TCVN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n_items * 128 * 128, st));
    return TCVN_OK;
  }
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one launch");
  WgradParams p{};
  p.rows = rows; p.n_items = n_items;
  for (int i = 0; i < n_items; ++i) { p.item_col[i] = item_col[i]; p.item_shift[i] = item_shift[i]; p.item_valid[i] = item_valid[i]; }
  p.a_cols = a_fold_cols; p.a_scale = a_scale; p.a_shift = a_shift; p.a_alpha = a_alpha;
  p.g_col0 = g_col0; p.dw = parts;
  p.num_tiles = (int)ceil_div_ll(rows, kWgRows);
  const int stage_bytes = (2 * n_items + 2) * kWgBox;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return fail(TCVN_ERR_UNSUPPORTED, "umma_wgrad: stage of %d bytes leaves no room to pipeline", stage_bytes);
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 26 * 8 + 16;
  bool& attr_done = device_flag(3);   // per device: the attribute belongs to the device's copy of the function
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TCVN_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  CUtensorMap tmA, tmG;
  TCVN_TRY(make_map(A, rows, a_cols, a_pitch, 64, kWgRows, &tmA));
  TCVN_TRY(make_map(G, rows, g_cols, g_pitch, 64, kWgRows, &tmG));
  // few tiles: fewer CTAs (each partial costs a 64 KB store + read per item); many tiles: one CTA per SM
  int grid = ceil_div(p.num_tiles, 4);
  if (grid > sm_count()) grid = sm_count();
  if (grid > kWgMaxCtas) grid = kWgMaxCtas;
  if (grid < 1) grid = 1;
  if (a_scale) umma_wgrad_kernel<true><<<grid, kWgThreads, smem, st>>>(tmA, tmG, p);
  else umma_wgrad_kernel<false><<<grid, kWgThreads, smem, st>>>(tmA, tmG, p);
  TCVN_LAUNCH_CHECK();
  const long long n4 = (long long)n_items * 128 * 128 / 4;
  if (map_mode != 0) {   // reduce the per-CTA partials straight into the reference-layout gradient tensor
    MapArgs ma{map_mode, map_n_log, map_k_log, map_c0, map_c0p};
    reduce_parts_map_kernel<<<(unsigned)ceil_div_ll(n4, 32), 256, 0, st>>>(reinterpret_cast<const float4*>(parts), grid, n4, ma, dw);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }
  reduce_parts_kernel<<<(unsigned)ceil_div_ll(n4, 32), 256, 0, st>>>(reinterpret_cast<const float4*>(parts), grid, n4,
                                                                     reinterpret_cast<float4*>(dw), accumulate ? 1 : 0);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// out[p][c] (bf16 [rows, 128], ring rows zero) = sum_dy sum_k G2x[p + (1 - dy) * Wp][k] * Wd[dy][c][k]
//   G2x  bf16 [rows, 128]: columns dx*32 + n = G[p + 1 - dx][n], columns 96.. zero
//   Wd   bf16 [3][128][128]: Wd[dy][c][dx*32 + n] = w2[n][c][dy][dx], columns 96.. zero
int umma_conv2_dgrad(const void* g2x, const void* wd, long long rows, int Hp, int Wp, void* out, cudaStream_t st,
                     const void* red_x, const float* red_fold, double* red_parts, int* red_slots) {
  if (red_slots) *red_slots = 0;
  if (rows <= 0) return TCVN_OK;
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one launch");
  DgradParams p{};
  p.m_total = rows; p.Hp = Hp; p.Wp = Wp;
  p.nbox = ceil_div(128 + 2 * Wp, kDgBoxRows);
  p.halo_rows = p.nbox * kDgBoxRows;
  const int halo_rows_max = 288;
  if (p.halo_rows > halo_rows_max)
    return fail(TCVN_ERR_UNSUPPORTED, "feature map width %d needs a %d-row halo tile (max %d)", Wp - 2, p.halo_rows, halo_rows_max);
  p.out = static_cast<bf16*>(out); p.ldo = 128;
  p.red_x = static_cast<const bf16*>(red_x); p.red_fold = red_fold; p.red_parts = red_x ? red_parts : nullptr;
  p.num_tiles = (int)ceil_div_ll(rows, 128);
  bool& attr_done = device_flag(4);
  const size_t smem_max = 1024 + kDgWBytes + (size_t)kDgStages * halo_rows_max * 128 + 256;
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_conv2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    attr_done = true;
  }
  CUtensorMap tmG, tmW;
  TCVN_TRY(make_map(g2x, rows, 128, 128, 64, kDgBoxRows, &tmG));
  TCVN_TRY(make_map(wd, 3 * 128, 128, 128, 64, 128, &tmW));
  const size_t smem = 1024 + kDgWBytes + (size_t)kDgStages * p.halo_rows * 128 + 256;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (red_slots) *red_slots = p.red_parts ? grid : 0;
  umma_conv2_dgrad_kernel<<<grid, kDgThreads, smem, st>>>(tmG, tmW, p);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn

// ---- C ABI: the two kernels on caller-provided matrices (unit-tested against plain matrix products) ----------
extern "C" size_t tcvn_t_umma_wgrad_workspace_bytes(int n_items) { return tcvn::umma_wgrad_parts_bytes(n_items); }

extern "C" int tcvn_t_umma_wgrad(const void* a_bf16, int64_t rows, int a_cols, int a_pitch, int n_items,
                                 const int32_t* item_col, const int32_t* item_shift, const int32_t* item_valid,
                                 const float* a_fold, int a_fold_cols, const void* g_bf16, int g_cols, int g_pitch, int g_col0,
                                 float* dw, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(a_bf16 && g_bf16 && dw && item_col && item_shift && item_valid && workspace, "t_umma_wgrad: null pointer");
  TCVN_CHECK_ARG(n_items >= 1 && n_items <= 4, "t_umma_wgrad: 1..4 items");
  if (workspace_bytes < tcvn::umma_wgrad_parts_bytes(n_items))
    return tcvn::fail(TCVN_ERR_WORKSPACE, "t_umma_wgrad: workspace %zu < %zu bytes", workspace_bytes, tcvn::umma_wgrad_parts_bytes(n_items));
  TCVN_CHECK_ARG(a_pitch % 8 == 0 && g_pitch % 8 == 0, "t_umma_wgrad: row pitches must be multiples of 8 elements (TMA)");
  TCVN_CHECK_ARG(a_fold == nullptr || a_fold_cols % 8 == 0, "t_umma_wgrad: fold width must be a multiple of 8");
  return tcvn::umma_wgrad(a_bf16, rows, a_cols, a_pitch, n_items, item_col, item_shift, item_valid, a_fold,
                          a_fold ? a_fold + a_fold_cols : nullptr, a_fold ? a_fold + 2 * a_fold_cols : nullptr, a_fold_cols,
                          g_bf16, g_cols, g_pitch, g_col0, static_cast<float*>(workspace), dw, true, stream);
}

extern "C" int tcvn_t_umma_conv2_dgrad(const void* g2x_bf16, const void* wd_bf16, int64_t rows, int ring_hp, int ring_wp,
                                       void* out_bf16, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(g2x_bf16 && wd_bf16 && out_bf16 && ring_hp >= 3 && ring_wp >= 3, "t_umma_conv2_dgrad: bad arguments");
  return tcvn::umma_conv2_dgrad(g2x_bf16, wd_bf16, rows, ring_hp, ring_wp, out_bf16, stream, nullptr, nullptr, nullptr, nullptr);
}
