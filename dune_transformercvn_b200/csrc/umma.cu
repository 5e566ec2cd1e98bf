// bf16 tcgen05 path of the DenseNet bottleneck layer (dense_net.py:8-45), sm_100a.
//
//   conv1 kernel   mid = PReLU2(BN2(conv1x1(PReLU1(BN1(blk[:, :k]))) + b1))
//       GEMM  M = pixels (128-row tiles of the ringed channels-last buffer), N = 128, K = k (64-wide chunks).
//       TMA loads the RAW concat tile and the weight chunk; 4 transform warps apply BN1+PReLU1 in place in
//       the 128B-swizzled tile (each layer has its own BN over the same concat channels, so the activation
//       cannot be stored once); one thread issues tcgen05.mma into a double-buffered TMEM accumulator;
//       4 epilogue warps read TMEM, apply bias+BN2+PReLU2, zero the ring rows and store bf16.
//   conv2 kernel   blk[:, k:k+32] = conv3x3(mid) + b2      (zero padding == the ring of `mid`)
//       shifted GEMM: a 3x3 tap is a constant row offset in the ringed layout, so ONE haloed tile
//       (128 + 2*(Wp+1) rows x 128 ch) is loaded per output tile and the 9 taps are 9 MMA groups whose A
//       descriptors start at different rows of that tile; the 9x32x128 weights stay resident in SMEM.
// Both kernels are persistent (one CTA per SM, static round-robin over tiles) and warp-specialised.
#include <mutex>
#include <map>
#include <tuple>
#include <stdlib.h>

#include "ptx.cuh"
#include "umma.h"

namespace tcvn {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 row-major matrix [rows][cols] with `pitch` elements per row; box = box_cols x box_rows, 128B swizzle,
// out-of-bounds elements (negative rows included) read as zero.
int make_map(const void* base, long long rows, int cols, int pitch, int box_cols, int box_rows, CUtensorMap* out) {
  // The only host-side state of the library besides the error text: a per-THREAD memo of encoded tensor maps (an encode
  // is ~1.5 us of driver time, an eager training step makes ~5 000 of them).  A map is a pure function of its key, so an
  // entry can never be stale; the memo is bounded and shared with nobody (no lock).  Under CUDA-graph replay - the default
  // of both hot paths - it is not consulted at all.
  typedef std::tuple<const void*, long long, int, int, int, int> Key;
  static thread_local std::map<Key, CUtensorMap> cache;
  Key key(base, rows, cols, pitch, box_cols, box_rows);
  {
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return TCVN_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(TCVN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // L2 promotion NONE: a 256-byte promotion fetches the neighbouring line of every 128-byte box row whose line is the lower
  // half of a 256-byte block - for the block buffers, where a layer reads the first K channels of wider rows, that was 445 MB
  // of DRAM reads for 356 MB needed (ncu).  For maps whose rows are whole blocks read half by half (mid and its gradients)
  // promotion would be a free prefetch; measured (scripts/gpu_round2_promo3.sh): no difference, so one setting for all.
  // TCVN_TMAP_PROMO = 128 / 256 forces another one (A/B).
  static const CUtensorMapL2promotion promo = [] {
    const char* v = getenv("TCVN_TMAP_PROMO");
    const int n = v ? atoi(v) : 0;
    return n == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : n == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : n == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  }();
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(TCVN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d pitch=%d box=%dx%d", (int)r, rows,
                cols, pitch, box_cols, box_rows);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return TCVN_OK;
}

// SMs the persistent tcgen05 kernels launched from THIS host thread may occupy (0 = all).  The two pixel-map CNNs run on
// two streams; with every kernel sized for the whole chip they only time-slice, with disjoint SM budgets they run side by
// side and each stream's launch prologues (TMEM allocation, barrier setup, 72 KB of conv2 weights: ~7 us per launch) hide
// under the other stream's kernels.
static thread_local int g_sm_limit = 0;

constexpr int kMaxDevices = 64;

int sm_count() {
  static int per_device[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = per_device[dev >= 0 && dev < kMaxDevices ? dev : 0];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return g_sm_limit > 0 && g_sm_limit < n ? g_sm_limit : n;
}

// one-time-per-device flags of the launchers (cudaFuncSetAttribute is per device); `site` = a launcher id in [0, 8)
bool& device_flag(int site) {
  static bool flags[kMaxDevices][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  return flags[dev][site & 7];
}

}  // namespace tcvn

extern "C" int tcvn_set_sm_limit(int n_sms) {
  if (n_sms < 0) return tcvn::fail(TCVN_ERR_ARG, "set_sm_limit: %d", n_sms);
  tcvn::g_sm_limit = n_sms;
  return TCVN_OK;
}

namespace tcvn {

// ------------------------------------------------------------------------------------------------
// GEMM with fused operand activation:  out[m, n] = PReLU_n( sum_k act_k(A[m, k]) * W[n, k] + shift[n] )
//   conv1 of a bottleneck: act = BN1+PReLU1 per input channel, W = BN2-scaled conv1 weights, shift = folded
//                          bias/BN2, PReLU2 slopes;  N = 128
//   transition conv:       act = identity (input already activated + pooled), shift = bias, slope 1; N tiles of 128
// Warp roles (576 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      TMA producer: raw A tile [128 rows x 64 ch] + W chunk [128 n x 64 k] per stage
//   warp 1      MMA issuer (one thread): 4 x tcgen05.mma (M128 N128 K16) per stage into TMEM accumulator acc
//   warps 2-9   operand transform in place in the 128B-swizzled A tile (generic proxy -> fence -> async proxy)
//   warps 10-17 two epilogue groups, one per TMEM accumulator: TMEM -> regs -> +shift, PReLU, ring rows -> 0,
//               bf16 -> swizzled staging tile -> TMA store
// ------------------------------------------------------------------------------------------------
constexpr int kC1Stages = 4;     // pipeline stages when the weights stream with the activations (one A + one W chunk per stage)
constexpr int kC1StagesMax = 8;  // ... and when the weights are resident (at most four (N tile, K chunk) slots): the 16 KB W slots they leave free become A stages
constexpr int kC1Acc = 4;       // TMEM accumulators (all 512 columns): the MMA warp runs up to two tiles ahead of each epilogue group
constexpr int kC1Threads = 576;
constexpr int kTileM = 128;
constexpr int kMid = 128;              // N tile = bottleneck width the kernels are specialised for
constexpr int kStageA = kTileM * 128;  // 16 KB: 128 rows x 64 bf16
constexpr int kStageW = kMid * 128;    // 16 KB: 128 out-channels x 64 bf16
constexpr int kXformThreads = 256;

struct GemmParams {
  long long m_total;
  int kchunks, kphys;      // K chunks of 64; channels >= kphys are forced to zero by the transform
  // shifted-GEMM form (3x3 convolutions of the --sdxl CNN): the K loop runs over n_taps row-shifted views of A
  // (chunks_per_tap chunks each, rows m + tap_off[t]: a 3x3 tap is a constant row offset in the ringed layout, rows outside
  // the matrix read as zero) followed by chunks2 chunks of a SECOND matrix (tmA2, no shift) - the residual input of a
  // ResNet block (identity or 1x1-shortcut weights in the matching K range of W).  Plain GEMM: n_taps = 1, tap_off = {0}.
  int n_taps, chunks_per_tap, chunks2;
  int tap_off[9];
  int tap_col[9];   // first column of every tap in A (0: all taps read the same columns)
  int ntile;               // output channels per N tile: 128, or 64 (the 64-channel layers of the --sdxl CNN: half the MMA
                           // work and half the weight traffic of a zero-padded 128-wide tile)
  int n_tiles_n;           // N tiles of 128 (1 for conv1)
  const float *a_scale, *a_shift, *a_alpha;  // [kchunks*64] (TRANSFORM only)
  const float *o_shift, *o_alpha;            // [n_tiles_n*128]
  int Hp, Wp, num_tiles;   // num_tiles = m tiles * n_tiles_n
  // training: per-column (sum, sum^2) of the bf16 values this kernel stores (ring rows are zero) - the batch statistics of
  // the BatchNorm that follows (n_tiles_n == 1).  Every (CTA, epilogue group) stores its partial sums in its own slot,
  // stats[(2 * cta + group)][2][128]; the consumer adds the slots in a fixed order (no atomics: bit-reproducible).
  double* stats;
  // weights-resident mode (N tiles x K chunks <= 4: every conv1 of dense blocks 1-2, the first transition, the conv1
  // input-gradient GEMMs up to K = 256): the whole weight matrix is loaded once per CTA, slot (nt, kc), and the pipeline has
  // `stages` A-only stages.  The clock-stamp trace of the streaming form showed a
  // stage cycle of 3.2 us (1.2-2.2 us of it TMA latency under load) with 4 stages in flight: 0.8 us per chunk, below what
  // HBM delivers - the kernel was bound by bytes in flight, not by bandwidth, the epilogue or the tensor pipe.
  int wres, stages;
};

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float prelu_fast(float v, float a) { return fmaf(a, fminf(v, 0.f), fmaxf(v, 0.f)); }

// MMASHIFT (single N tile): the epilogue's "+ shift[n]" is done by the tensor core.  One extra K16 MMA per tile multiplies a
// constant A tile (k0 = k1 = 1) with a B tile holding shift[n] split into two bf16 (k0 = hi, k1 = shift - hi: 16 mantissa
// bits, products with 1.0 are exact), issued FIRST so that it also initialises the accumulator.  ncu showed the kernel
// bound by the shared-memory data pipe (79 % LSU wavefronts), a third of them the epilogue's broadcast loads of the
// per-column constants: this removes the 256 shift wavefronts per tile, and the slopes are read 8 columns per load.
template <bool TRANSFORM, bool MMASHIFT>
__global__ void __launch_bounds__(kC1Threads, 1) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW,
                                                                  const __grid_constant__ CUtensorMap tmO,
                                                                  const __grid_constant__ CUtensorMap tmA2,
                                                                  const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // [stages][16 KB]
  uint8_t* sW = smem + p.stages * kStageA;              // [stages][16 KB], or [kchunks][16 KB] resident
  uint8_t* sOut = sW + (p.wres ? p.kchunks * p.n_tiles_n : p.stages) * kStageW;   // [2 groups][2 halves][128 rows x 128 B], swizzled
  uint8_t* sXA = sOut + 4 * kStageA;                    // [128 rows x 128 B]: k0 = k1 = 1.0, rest 0   (MMASHIFT)
  uint8_t* sXB = sXA + kStageA;                         // [128 n x 128 B]: k0 = hi(shift[n]), k1 = lo(shift[n])
  float* s_epi = reinterpret_cast<float*>(sXB + kStageW);  // [2 groups][128 shift + 64 packed slopes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epi + 2 * 192);
  uint64_t* full = bars;                    // TMA -> transform (or MMA)
  uint64_t* ready = bars + kC1StagesMax;       // transform -> MMA
  uint64_t* empty = bars + 2 * kC1StagesMax;   // MMA -> TMA
  uint64_t* tfull = bars + 3 * kC1StagesMax;   // MMA -> epilogue   [kC1Acc]
  uint64_t* tempty = tfull + kC1Acc;        // epilogue -> MMA   [kC1Acc]
  uint64_t* wfull = tempty + kC1Acc;           // resident weight slot (nt, kc) has landed   [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int kc = 0; kc < 4; ++kc) ptx::mbar_init(&wfull[kc], 1);
    for (int s = 0; s < kC1StagesMax; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&ready[s], kXformThreads);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < kC1Acc; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmO);
    ptx::prefetch_tmap(&tmA2);
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, kC1Acc * kMid);
  if (MMASHIFT) {
    uint4* z = reinterpret_cast<uint4*>(sXA);
    for (int i = threadIdx.x; i < (kStageA + kStageW) / 16; i += kC1Threads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < kTileM) {   // row r: elements k0, k1 live in 16-byte chunk 0, stored at chunk position (r & 7)
      const int r = threadIdx.x;
      const float sv = __ldg(p.o_shift + r);
      const bf16 hi = __float2bfloat16_rn(sv);
      const bf16 lo = __float2bfloat16_rn(sv - __bfloat162float(hi));
      *reinterpret_cast<uint32_t*>(sXA + r * 128 + ((r & 7) << 4)) = 0x3F803F80u;
      *reinterpret_cast<uint32_t*>(sXB + r * 128 + ((r & 7) << 4)) =
          (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (p.wres) {   // one barrier per slot: the first tile's MMAs start as soon as THEIR chunk is there
        for (int n = 0; n < p.n_tiles_n; ++n)
          for (int kc = 0; kc < p.kchunks; ++kc) {
            const int slot = n * p.kchunks + kc;
            ptx::mbar_arrive_expect_tx(&wfull[slot], p.ntile * 128);
            ptx::tma_load_2d(sW + slot * kStageW, &tmW, &wfull[slot], kc * 64, n * p.ntile);
          }
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
        const int tap_chunks = p.n_taps * p.chunks_per_tap;
        int tap = 0, kk = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage], kStageA + (p.wres ? 0 : p.ntile * 128));
          if (kc < tap_chunks) {
            ptx::tma_load_2d(sA + stage * kStageA, &tmA, &full[stage], kk * 64 + p.tap_col[tap], mt * kTileM + p.tap_off[tap]);
            if (++kk == p.chunks_per_tap) { kk = 0; ++tap; }
          } else {
            ptx::tma_load_2d(sA + stage * kStageA, &tmA2, &full[stage], (kc - tap_chunks) * 64, mt * kTileM);
          }
          if (!p.wres) ptx::tma_load_2d(sW + stage * kStageW, &tmW, &full[stage], kc * 64, nt * p.ntile);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      // a CTA does not visit every N tile (gridDim.x is a multiple of n_tiles_n more often than not), so the MMA thread may
      // never wait for some resident weight slots: no bulk copy into this CTA's shared memory may be in flight when it exits
      if (p.wres)
        for (int slot = 0; slot < p.n_tiles_n * p.kchunks; ++slot) ptx::mbar_wait(&wfull[slot], 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(kTileM, p.ntile);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      uint32_t w_seen = p.wres ? 0u : 0xffu;   // resident weight slots this thread has already waited for
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMid;
        if (MMASHIFT)   // accumulator := 1 * shift[n]
          ptx::umma_bf16(d_tmem, ptx::umma_desc_join(ptx::kUmmaDescHiSw128, ptx::umma_desc_lo(ptx::smem_u32(sXA))),
                         ptx::umma_desc_join(ptx::kUmmaDescHiSw128, ptx::umma_desc_lo(ptx::smem_u32(sXB))), idesc, 0);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          ptx::mbar_wait(TRANSFORM ? &ready[stage] : &full[stage], phase);
          const int wslot = p.wres ? (tile % p.n_tiles_n) * p.kchunks + kc : stage;
          if (!((w_seen >> wslot) & 1u)) { ptx::mbar_wait(&wfull[wslot], 0); w_seen |= 1u << wslot; }
          ptx::tc_fence_after();
          const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(sA + stage * kStageA));
          const uint32_t w_lo = ptx::umma_desc_lo(ptx::smem_u32(sW + wslot * kStageW));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(d_tmem, ptx::umma_desc_join(ptx::kUmmaDescHiSw128, a_lo + 2 * k),
                           ptx::umma_desc_join(ptx::kUmmaDescHiSw128, w_lo + 2 * k), idesc, MMASHIFT || (kc | k) != 0);
          ptx::umma_commit(&empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull[acc]);
        if (++acc == kC1Acc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 10) {
    if (TRANSFORM) {
      // warp w owns channel group cg = w (8 channels of the K chunk) for all 128 rows, lane l the rows l, l+32, l+64,
      // l+96.  With the 128B swizzle those channels sit at 16-byte position cg ^ (row & 7) of the row, and
      // (l + 32 i) & 7 == l & 7: a thread always touches the same position, its constants stay in registers, a warp's
      // 32 x 16 B access covers every bank four times (4 wavefronts per 512 B: the minimum), and - unlike a mapping
      // with all 8 groups in one warp - the constant loads are warp-wide broadcasts.
      const int t = threadIdx.x - 64;
      const int cg = t >> 5;
      const int rb = lane, pc = cg ^ (lane & 7);
      int stage = 0; uint32_t phase = 0;
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      // the 8 channels' constants: BN in fp32, PReLU on packed bf16 pairs (max/min are exact; one rounding of slope *
      // negative part)
      struct Consts { float sc[8], sh[8]; __nv_bfloat162 al2[4]; bool live; };
      auto load_consts = [&](int kc, Consts& c) {
        const int ch = kc * 64 + cg * 8;
        c.live = ch < p.kphys;  // kphys is a multiple of 8: a group is all-live or all-dead
        const float4* s4 = reinterpret_cast<const float4*>(p.a_scale + ch);
        const float4* h4 = reinterpret_cast<const float4*>(p.a_shift + ch);
        const float4* a4 = reinterpret_cast<const float4*>(p.a_alpha + ch);
        const float4 x0 = __ldg(s4), x1 = __ldg(s4 + 1), y0 = __ldg(h4), y1 = __ldg(h4 + 1), z0 = __ldg(a4), z1 = __ldg(a4 + 1);
        c.sc[0] = x0.x; c.sc[1] = x0.y; c.sc[2] = x0.z; c.sc[3] = x0.w; c.sc[4] = x1.x; c.sc[5] = x1.y; c.sc[6] = x1.z; c.sc[7] = x1.w;
        c.sh[0] = y0.x; c.sh[1] = y0.y; c.sh[2] = y0.z; c.sh[3] = y0.w; c.sh[4] = y1.x; c.sh[5] = y1.y; c.sh[6] = y1.z; c.sh[7] = y1.w;
        c.al2[0] = __floats2bfloat162_rn(z0.x, z0.y); c.al2[1] = __floats2bfloat162_rn(z0.z, z0.w);
        c.al2[2] = __floats2bfloat162_rn(z1.x, z1.y); c.al2[3] = __floats2bfloat162_rn(z1.z, z1.w);
      };
      auto do_chunk = [&](const Consts& c) {
        if (lane == 0) ptx::mbar_wait(&full[stage], phase);  // one poller per warp keeps the LSU free
        __syncwarp();
        uint8_t* base = sA + stage * kStageA + rb * 128 + pc * 16;
        if (c.live) {
          uint4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const uint4*>(base + i * 32 * 128);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __nv_bfloat162 y = __floats2bfloat162_rn(fmaf(bf_lo(w[q]), c.sc[2 * q], c.sh[2 * q]),
                                                             fmaf(bf_hi(w[q]), c.sc[2 * q + 1], c.sh[2 * q + 1]));
              const __nv_bfloat162 r = __hfma2(c.al2[q], __hmin2(y, zero2), __hmax2(y, zero2));
              w[q] = *reinterpret_cast<const uint32_t*>(&r);
            }
            *reinterpret_cast<uint4*>(base + i * 32 * 128) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(base + i * 32 * 128) = make_uint4(0u, 0u, 0u, 0u);
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&ready[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      };
      if (p.kchunks <= 2) {
        // K <= 128 (dense block 1, the most expensive layers): the constants of both K chunks stay in registers for
        // every tile of this CTA.  ncu: the per-chunk reloads were 40 % of the kernel's LSU wavefronts (35 k of 87 k per
        // SM), more than the operand transform itself.
        Consts c0, c1;
        load_consts(0, c0);
        load_consts(p.kchunks - 1, c1);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          do_chunk(c0);
          if (p.kchunks == 2) do_chunk(c1);
        }
      } else {
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            Consts c;
            load_consts(kc, c);
            do_chunk(c);
          }
        }
      }
    }
  } else {
    const int grp = (warp - 10) >> 2;   // epilogue group == TMEM accumulator it drains
    const int g = warp & 3;             // TMEM lane group this warp may read
    const int row = g * 32 + lane;
    const int R = p.Hp * p.Wp;
    const bool issuer = threadIdx.x == (10 + 4 * grp) * 32;
    uint8_t* stg = sOut + grp * 2 * kStageA;
    float* s_shift = s_epi + grp * 192;                                   // [128] fp32
    uint32_t* s_alpha2 = reinterpret_cast<uint32_t*>(s_shift + kMid);     // [64] bf16x2
    int it = 0;
    int mine = 0;   // tiles this group has drained: accumulator grp + 2 * (mine & 1), barrier parity (mine >> 1) & 1
    float st1[8], st2[8];   // training statistics: this thread's 8 columns over its rows of every tile (fp32 partials)
#pragma unroll
    for (int q = 0; q < 8; ++q) { st1[q] = 0.f; st2[q] = 0.f; }
    const int e_col = threadIdx.x - (10 + 4 * grp) * 32;
    // epilogue constants of N tile nt -> shared memory (shift fp32, PReLU slopes as packed bf16 pairs).  With a single N
    // tile (every conv1) they are loaded once; per tile, their L2 latency sat in front of every tile of an epilogue-bound
    // group (the K = 64 layers).
    auto load_consts = [&](int nt) {
      if (!MMASHIFT) s_shift[e_col] = __ldg(p.o_shift + nt * p.ntile + e_col);
      if (e_col < kMid / 2) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(p.o_alpha + nt * p.ntile) + e_col);
        const __nv_bfloat162 a2 = __floats2bfloat162_rn(a.x, a.y);
        s_alpha2[e_col] = *reinterpret_cast<const uint32_t*>(&a2);
      }
    };
    if (p.n_tiles_n == 1) load_consts(0);   // visible to the group after the first named barrier below
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
      const long long m = (long long)mt * kTileM + row;
      bool ring = false;
      {
        const uint32_t rr = (uint32_t)m % (uint32_t)R;   // m < 2^31 (checked by the launcher): 32-bit division
        const uint32_t y = rr / (uint32_t)p.Wp, x = rr - y * (uint32_t)p.Wp;
        ring = y == 0 || y == (uint32_t)p.Hp - 1 || x == 0 || x == (uint32_t)p.Wp - 1;
      }
      if (p.n_tiles_n != 1) load_consts(nt);
      const int acc = grp + 2 * (mine & 1);
      const uint32_t acc_phase = (uint32_t)(mine >> 1) & 1u;
      ++mine;
      if (lane == 0) ptx::mbar_wait(&tfull[acc], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      if (issuer) ptx::tma_store_wait_read();  // this group's previous store has drained the staging tile
      ptx::named_bar_sync(1 + grp, 128);
#pragma unroll 1
      for (int c = 0; c < (p.ntile >> 5); c += 2) {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + acc * kMid + c * 32, r0);
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + acc * kMid + c * 32 + 32, r1);
        ptx::tmem_ld_wait();
        uint8_t* orow = stg + (c >> 1) * kStageA + row * 128;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t* r = hh ? r1 : r0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t o[4];
            const uint4 a4 = *reinterpret_cast<const uint4*>(s_alpha2 + (((c + hh) * 32 + q * 8) >> 1));   // slopes of 8 columns
            const uint32_t a2w[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int n = (c + hh) * 32 + q * 8 + h * 4;
              float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
              if (!MMASHIFT) sh = *reinterpret_cast<const float4*>(s_shift + n);
              const int j = q * 8 + h * 4;
              const __nv_bfloat162 y0 = __floats2bfloat162_rn(__uint_as_float(r[j + 0]) + sh.x, __uint_as_float(r[j + 1]) + sh.y);
              const __nv_bfloat162 y1 = __floats2bfloat162_rn(__uint_as_float(r[j + 2]) + sh.z, __uint_as_float(r[j + 3]) + sh.w);
              const __nv_bfloat162 p0 = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&a2w[2 * h]), __hmin2(y0, zero2), __hmax2(y0, zero2));
              const __nv_bfloat162 p1 = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&a2w[2 * h + 1]), __hmin2(y1, zero2), __hmax2(y1, zero2));
              o[2 * h] = *reinterpret_cast<const uint32_t*>(&p0);
              o[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&p1);
            }
            const int chunk = hh * 4 + q;  // 16-byte chunk within the 128-byte half-row
            *reinterpret_cast<uint4*>(orow + ((chunk ^ (row & 7)) << 4)) =
                ring ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[acc]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1 + grp, 128);
      if (issuer) {
        ptx::tma_store_2d(&tmO, stg, nt * p.ntile, mt * kTileM);
        if (p.ntile > 64) ptx::tma_store_2d(&tmO, stg + kStageA, nt * p.ntile + 64, mt * kTileM);
        ptx::tma_store_commit();
      }
      if (MMASHIFT && p.stats != nullptr) {   // (statistics need the single-N-tile form, which is the MMASHIFT one: the other
                                              // instantiations drop the accumulators - they spilled registers for nothing)
        // statistics of exactly the bf16 values being stored, read back from the staged tile 8 columns (16 bytes) at a
        // time: thread e owns column group e & 15 on rows (e >> 4) + 8 i.  (One column per thread cost 128 two-byte
        // loads per tile - as many LSU wavefronts as the operand transform - and made the train-mode kernel 35 % slower
        // than the eval one.)  Rows beyond m_total hold shift-only values and are skipped explicitly.
        const int sg = e_col & 15, r0 = e_col >> 4;
        const uint8_t* gp = stg + (sg >> 3) * kStageA;
        const long long m0 = (long long)mt * kTileM;
        const int rmax = (int)(p.m_total - m0 < kTileM ? p.m_total - m0 : kTileM);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int r = r0 + 8 * i;
          if (r >= rmax) break;
          const uint4 v = *reinterpret_cast<const uint4*>(gp + r * 128 + (((sg & 7) ^ (r & 7)) << 4));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = bf_lo(w[q]), b = bf_hi(w[q]);
            st1[2 * q] += a; st1[2 * q + 1] += b;
            st2[2 * q] = fmaf(a, a, st2[2 * q]); st2[2 * q + 1] = fmaf(b, b, st2[2 * q + 1]);
          }
        }
      }
    }
    if (issuer) ptx::tma_store_wait_all();
    if (MMASHIFT && p.stats != nullptr) {
      // per-group reduction through the (now idle) staging tile in a fixed order, then one store per column into the slot
      ptx::named_bar_sync(1 + grp, 128);          // the issuer's stores have drained the staging tile
      float* red = reinterpret_cast<float*>(stg);  // [128 threads][16]
#pragma unroll
      for (int q = 0; q < 8; ++q) { red[e_col * 16 + q] = st1[q]; red[e_col * 16 + 8 + q] = st2[q]; }
      ptx::named_bar_sync(1 + grp, 128);
      // column c = 8 * group + q  <-  threads e with (e & 15) == group, e = group + 16 j
      const int cgp = e_col >> 3, cq = e_col & 7;
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a += (double)red[(cgp + 16 * j) * 16 + cq];
        b += (double)red[(cgp + 16 * j) * 16 + 8 + cq];
      }
      double* slot = p.stats + (size_t)(2 * blockIdx.x + grp) * 2 * kMid;   // zeros when this group drained no tile
      slot[e_col] = a;
      slot[kMid + e_col] = b;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, kC1Acc * kMid);
}

// ------------------------------------------------------------------------------------------------
// conv2: 3x3 convolution as a shifted GEMM over one haloed tile, vertical taps in the MMA, horizontal
// taps in the epilogue.
//   Y[q][(dx, n)] = sum_dy sum_c mid[q + (dy-1)*Wp][c] * W[dy][dx][n][c]      3 x 8 MMAs of M128 N96 K16
//   out[p][n]     = Y[p-1][(0,n)] + Y[p][(1,n)] + Y[p+1][(2,n)] + b[n]        lane shuffles after tcgen05.ld
// An MMA with N = 32 spends its time re-reading the 4 KB A slab from shared memory (measured: tensor pipe
// 74 % busy at 28 % of peak); widening N to the three horizontal taps reads A a third as often.  Row p+-1
// is the neighbouring TMEM lane, so each tile produces 126 output rows from 128 Y rows; the two lanes at
// every warp boundary are exchanged through shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kC2StagesMax = 4;   // haloed tiles in flight: as many as fit beside the 72 KB of weights (2 on the 99x69 maps, 3 from block 3 on)
constexpr int kC2Threads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 two epilogue groups (one per TMEM accumulator)
constexpr int kGrowth = 32;
constexpr int kBoxRows = 32;
constexpr int kC2N = 3 * kGrowth;               // 96 accumulator columns: (dx, n)
constexpr int kC2Out = kTileM - 2;              // 126 output rows per tile
constexpr int kW2Slab = kC2N * 128;             // 12 KB: [96 rows x 128 B] per (dy, half)
constexpr int kW2Bytes = 3 * 2 * kW2Slab;       // 72 KB
constexpr size_t kC2SmemMax = 227 * 1024;        // dynamic shared memory a CTA may ask for on sm_100

struct Conv2Params {
  long long m_total;
  int Hp, Wp, halo_rows, nbox;  // halo_rows = nbox * kBoxRows >= 128 + 2*Wp
  int stages;                   // haloed tiles in flight (2 .. kC2StagesMax)
  const float* bias;
  bf16* out;
  int ldo, col0, num_tiles;
  // training: Dropout(p) on the 32 new channels (mask = hash(seed, site, row * 32 + channel), re-derived in backward) and
  // their per-column (sum, sum^2) as stored (bf16): the CTA's eight epilogue warps are added in a fixed order and stored in
  // the CTA's own slot, stats[cta][2][32]; the consumer adds the slots in a fixed order (no atomics)
  float p_drop; unsigned long long seed, site; const unsigned long long* seed_off;
  double* stats;
};

__device__ __forceinline__ bool drop_keep_c2(unsigned long long seed, unsigned long long site, unsigned long long idx, float p) {
  unsigned long long z = seed * 0x100000001b3ull + site * 0x9e3779b97f4a7c15ull + idx;   // same hash as train.cu
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  const uint32_t r = (uint32_t)((z ^ (z >> 31)) >> 32);
  return (r >> 8) * (1.0f / 16777216.0f) >= p;
}

__global__ void __launch_bounds__(kC2Threads, 1) umma_conv2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmW,
                                                                   const Conv2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int half_bytes = p.halo_rows * 128;
  const int stage_bytes = 2 * half_bytes;
  uint8_t* sW = smem;
  uint8_t* sA = smem + kW2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + p.stages * stage_bytes);
  // the pipeline unit is HALF a haloed tile (64 of the 128 channels, <= 36 KB): four units in flight in the space of two
  // tiles, a unit is refilled as soon as its 12 MMAs have completed - the kernel is bound by the latency of getting a
  // haloed tile into shared memory, not by the tensor pipe
  uint64_t* full = bars;                      // [stage][half]
  uint64_t* empty = bars + 2 * kC2StagesMax;  // [stage][half]
  uint64_t* tfull = bars + 4 * kC2StagesMax;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_exch = s_bias + kGrowth;  // [2 groups][4 warps][2][32]: lane 31's dx=0 block, lane 0's dx=2 block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::mbar_init(wfull, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (threadIdx.x < kGrowth) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wfull, kW2Bytes);
      for (int dy = 0; dy < 3; ++dy)
        for (int h = 0; h < 2; ++h)
          ptx::tma_load_2d(sW + (dy * 2 + h) * kW2Slab, &tmW, wfull, h * 64, dy * kC2N);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row0 = tile * kC2Out - 1 - p.Wp;  // Y row 0 of the tile is output row tile*126 - 1
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(&empty[stage * 2 + h], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage * 2 + h], half_bytes);
          for (int b = 0; b < p.nbox; ++b)
            ptx::tma_load_2d(sA + stage * stage_bytes + h * half_bytes + b * kBoxRows * 128, &tmA, &full[stage * 2 + h], h * 64,
                             row0 + b * kBoxRows);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(kTileM, kC2N);
      ptx::mbar_wait(wfull, 0);
      const uint32_t w_lo = ptx::umma_desc_lo(ptx::smem_u32(sW));
      const uint32_t half16 = (uint32_t)half_bytes >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        const uint32_t d_tmem = tmem_base + acc * 128;
        // descriptors differ only in their 14-bit start-address field, in 16-byte units: +8 per row of the
        // haloed tile (any row is a legal start: the 128B swizzle is a function of the absolute address),
        // +2 per K step of 16 channels
        const uint32_t a_lo = ptx::umma_desc_lo(ptx::smem_u32(sA + stage * stage_bytes));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(&full[stage * 2 + h], phase);
          ptx::tc_fence_after();
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t row_lo = a_lo + (uint32_t)(dy * p.Wp) * 8u + h * half16;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t al = row_lo + k4 * 2u;
              const uint32_t bl = w_lo + (dy * 2 + h) * (kW2Slab >> 4) + k4 * 2u;
              ptx::umma_bf16(d_tmem, ptx::umma_desc_join(ptx::kUmmaDescHiSw128, al),
                             ptx::umma_desc_join(ptx::kUmmaDescHiSw128, bl), idesc, (h | dy | k4) != 0);
            }
          }
          ptx::umma_commit(&empty[stage * 2 + h]);
        }
        ptx::umma_commit(&tfull[acc]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    // two epilogue groups of four warps, group == the TMEM accumulator it drains (tiles alternate): ncu showed the
    // single group (tcgen05.ld, 64 lane shuffles, pack, store per tile) as the kernel's bottleneck - the MMA warp
    // waited for a free accumulator with the tensor pipe 44 % active
    const int grp = (warp - 2) >> 2;
    const int g = warp & 3;
    const int row = g * 32 + lane;
    const int R = p.Hp * p.Wp;
    float* s_exch_g = s_exch + grp * 256;
    float* ex_mine = s_exch_g + g * 64;
    const int acc = grp; uint32_t acc_phase = 0;
    double st_sum = 0.0, st_sq = 0.0;   // training statistics of column `lane`
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const long long m = (long long)tile * kC2Out + row - 1;  // output row of this lane
      bool ring = false;
      {
        const int rr = (int)((m + R) % R);
        const int y = rr / p.Wp, x = rr - y * p.Wp;
        ring = y == 0 || y == p.Hp - 1 || x == 0 || x == p.Wp - 1;
      }
      if (lane == 0) ptx::mbar_wait(&tfull[acc], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(g * 32) << 16) + acc * 128;
      uint32_t left[32], mid[32], right[32];
      ptx::tmem_ld_32x32(t_addr, left);
      ptx::tmem_ld_32x32(t_addr + 32, mid);
      ptx::tmem_ld_32x32(t_addr + 64, right);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[acc]);
      // publish the boundary lanes, then pull the neighbours' values: Y[p-1] block 0, Y[p+1] block 2
      ptx::named_bar_sync(3 + grp, 128);  // previous tile's exchange fully consumed
      if (lane == 31) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ex_mine[j] = __uint_as_float(left[j]);
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ex_mine[32 + j] = __uint_as_float(right[j]);
      }
      ptx::named_bar_sync(3 + grp, 128);
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float l = __shfl_up_sync(0xffffffffu, __uint_as_float(left[j]), 1);
        float r = __shfl_down_sync(0xffffffffu, __uint_as_float(right[j]), 1);
        if (lane == 0 && g > 0) l = s_exch_g[(g - 1) * 64 + j];
        if (lane == 31 && g < 3) r = s_exch_g[(g + 1) * 64 + 32 + j];
        o[j] = l + __uint_as_float(mid[j]) + r + s_bias[j];
      }
      const bool row_ok = row >= 1 && row <= kC2Out && m < p.m_total;
      if (p.p_drop > 0.f && row_ok && !ring) {
        const float inv = 1.f / (1.f - p.p_drop);
        const unsigned long long seed = seed_with_offset(p.seed, p.seed_off);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          o[j] = drop_keep_c2(seed, p.site, (unsigned long long)m * 32 + j, p.p_drop) ? o[j] * inv : 0.f;
      }
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = (ring || !row_ok) ? 0u : pack_bf16(o[2 * j], o[2 * j + 1]);
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + m * p.ldo + p.col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      }
      if (p.stats != nullptr) {   // warp-uniform
        float v1[32], v2[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v1[2 * j] = bf_lo(w[j]); v1[2 * j + 1] = bf_hi(w[j]);
          v2[2 * j] = v1[2 * j] * v1[2 * j]; v2[2 * j + 1] = v1[2 * j + 1] * v1[2 * j + 1];
        }
        st_sum += (double)warp_transpose_sum(v1, lane);
        st_sq += (double)warp_transpose_sum(v2, lane);
      }
      acc_phase ^= 1;
    }
    if (p.stats != nullptr) {
      // s_exch is idle by now (every tile's exchange has been consumed): [8 warps][2][32] doubles fit in its 2 KB
      double* red = reinterpret_cast<double*>(s_exch);
      ptx::named_bar_sync(5, 256);
      red[((4 * grp + g) * 2) * kGrowth + lane] = st_sum;
      red[((4 * grp + g) * 2 + 1) * kGrowth + lane] = st_sq;
      ptx::named_bar_sync(5, 256);
      if (grp == 0 && g < 2) {   // warp g reduces sum kind g over the eight warps in order
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[(w * 2 + g) * kGrowth + lane];
        p.stats[((size_t)blockIdx.x * 2 + g) * kGrowth + lane] = t;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline const float* pf(const char* packed, size_t off) { return reinterpret_cast<const float*>(packed + off); }

// dynamic shared memory of umma_gemm_kernel: 8 x 16 KB of operand slots (4 A + 4 W stages, or resident W + up to 8 A stages),
// 64 KB of output staging, the shift-MMA operands, epilogue constants, barriers
static size_t gemm_smem_bytes() {
  return 1024 + kC1Stages * (kStageA + kStageW) + 4 * kStageA + kStageA + kStageW + 2 * 192 * 4 +
         (3 * kC1StagesMax + 2 * kC1Acc + 4) * 8 + 16;
}
// weights-resident mode: at most four 16 KB weight slots (N tiles x K chunks); the A stages take the slots the weights leave free
static void gemm_pipeline(GemmParams& g) {
  static const bool on = [] { const char* v = getenv("TCVN_C1_WRES"); return !(v && v[0] == '0'); }();
  g.wres = on && g.kchunks * g.n_tiles_n <= 4 ? 1 : 0;
  g.stages = kC1Stages;
  if (g.wres) {
    const int slots = 2 * kC1Stages - g.kchunks * g.n_tiles_n;
    g.stages = slots < kC1StagesMax ? slots : kC1StagesMax;
  }
}

int launch_gemm(bool transform, const void* A, long long rows, int a_cols, int a_pitch, const void* W, int w_rows,
                       int kpad, int kphys, const float* a_scale, const float* a_shift, const float* a_alpha,
                       const float* o_shift, const float* o_alpha, void* out, int out_cols, int out_pitch, int n_tiles_n,
                       int Hp, int Wp, cudaStream_t st, double* stats, int* stat_slots) {
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one chunk");
  if (stats != nullptr && n_tiles_n != 1) return fail(TCVN_ERR_UNSUPPORTED, "epilogue statistics need a single N tile");
  if (stats != nullptr && getenv("TCVN_MMA_SHIFT") && getenv("TCVN_MMA_SHIFT")[0] == '0')
    return fail(TCVN_ERR_UNSUPPORTED, "epilogue statistics live in the shift-MMA form of the kernel (unset TCVN_MMA_SHIFT=0)");
  const size_t smem = gemm_smem_bytes();
  bool& attr_done = device_flag(1);   // per device: the attribute belongs to the device's copy of the function
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  CUtensorMap tmA, tmW, tmO;
  TCVN_TRY(make_map(A, rows, a_cols, a_pitch, 64, kTileM, &tmA));
  TCVN_TRY(make_map(W, w_rows, kpad, kpad, 64, kMid, &tmW));
  TCVN_TRY(make_map(out, rows, out_cols, out_pitch, 64, kTileM, &tmO));
  GemmParams g;
  g.m_total = rows; g.kchunks = kpad / kKChunk; g.kphys = kphys; g.n_tiles_n = n_tiles_n;
  g.n_taps = 1; g.chunks_per_tap = g.kchunks; g.chunks2 = 0; g.ntile = kMid;
  for (int t = 0; t < 9; ++t) g.tap_col[t] = 0;
  for (int t = 0; t < 9; ++t) g.tap_off[t] = 0;
  g.a_scale = a_scale; g.a_shift = a_shift; g.a_alpha = a_alpha; g.o_shift = o_shift; g.o_alpha = o_alpha;
  g.Hp = Hp; g.Wp = Wp;
  g.stats = stats;
  g.num_tiles = (int)ceil_div_ll(rows, kTileM) * n_tiles_n;
  gemm_pipeline(g);
  const int grid = g.num_tiles < sm_count() ? g.num_tiles : sm_count();
  if (stat_slots) *stat_slots = 2 * grid;
  static const bool mma_shift_on = [] { const char* v = getenv("TCVN_MMA_SHIFT"); return !(v && v[0] == '0'); }();
  const bool ms = mma_shift_on && n_tiles_n == 1;
  if (transform && ms) umma_gemm_kernel<true, true><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmA, g);
  else if (transform) umma_gemm_kernel<true, false><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmA, g);
  else if (ms) umma_gemm_kernel<false, true><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmA, g);
  else umma_gemm_kernel<false, false><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmA, g);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// Shifted GEMM on tcgen05 (3x3 / 1x1 convolutions of the --sdxl CNN over ringed channels-last bf16 maps):
//   out[m, n] = sum_t sum_c A[m + tap_off[t], c] * W[n, t * a_cols + c]  +  sum_c X2[m, c] * W[n, n_taps * a_cols + c]  +  bias[n]
// ring rows of `out` are written as zeros.  a_cols, x2_cols: multiples of 64; W bf16 [n_tiles * 128][n_taps * a_cols + x2_cols]
// (K-major, zero rows beyond the real output width); X2 may be null (x2_cols = 0).
int launch_gemm_shifted(const void* A, long long rows, int a_cols, int n_taps, const int* tap_off, const void* X2, int x2_cols,
                        const void* W, int n_tiles_n, const float* bias, const float* ones, void* out, int out_cols, int Hp, int Wp,
                        cudaStream_t st, int a_pitch, const int* tap_col) {
  if (a_pitch <= 0) a_pitch = a_cols;
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one chunk");
  if (a_cols % 64 || x2_cols % 64 || n_taps < 1 || n_taps > 9) return fail(TCVN_ERR_ARG, "launch_gemm_shifted: bad shape");
  const size_t smem = gemm_smem_bytes();
  bool& attr_done = device_flag(1);
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TCVN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int ktot = n_taps * a_cols + x2_cols;
  const int ntile = (n_tiles_n == 1 && out_cols <= 64) ? 64 : kMid;
  CUtensorMap tmA, tmW, tmO, tmX;
  TCVN_TRY(make_map(A, rows, a_pitch, a_pitch, 64, kTileM, &tmA));
  TCVN_TRY(make_map(W, (long long)n_tiles_n * kMid, ktot, ktot, 64, ntile, &tmW));
  TCVN_TRY(make_map(out, rows, out_cols, out_cols, 64, kTileM, &tmO));
  if (X2) TCVN_TRY(make_map(X2, rows, x2_cols, x2_cols, 64, kTileM, &tmX));
  else tmX = tmA;
  GemmParams g;
  g.m_total = rows; g.kchunks = ktot / kKChunk; g.kphys = ktot; g.n_tiles_n = n_tiles_n;
  g.n_taps = n_taps; g.chunks_per_tap = a_cols / kKChunk; g.chunks2 = x2_cols / kKChunk; g.ntile = ntile;
  for (int t = 0; t < 9; ++t) g.tap_off[t] = t < n_taps && tap_off ? tap_off[t] : 0;
  for (int t = 0; t < 9; ++t) g.tap_col[t] = t < n_taps && tap_col ? tap_col[t] : 0;
  g.a_scale = g.a_shift = g.a_alpha = nullptr; g.o_shift = bias; g.o_alpha = ones;
  g.Hp = Hp; g.Wp = Wp; g.stats = nullptr;
  g.num_tiles = (int)ceil_div_ll(rows, kTileM) * n_tiles_n;
  gemm_pipeline(g);
  const int grid = g.num_tiles < sm_count() ? g.num_tiles : sm_count();
  if (n_tiles_n == 1) umma_gemm_kernel<false, true><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmX, g);
  else umma_gemm_kernel<false, false><<<grid, kC1Threads, smem, st>>>(tmA, tmW, tmO, tmX, g);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int umma_dense_layer(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* pk, void* blk, void* mid,
                     long long rows, cudaStream_t st) {
  return umma_dense_layer_part(P, B, L, pk, blk, mid, rows, 3, st);
}

// which: bit 0 = conv1 (fused-activation GEMM), bit 1 = conv2 (3x3 shifted GEMM)
int umma_dense_layer_part(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* pk, void* blk, void* mid,
                          long long rows, int which, cudaStream_t st) {
  if (P.mid != kMid || P.d.growth != kGrowth)
    return fail(TCVN_ERR_UNSUPPORTED, "tcgen05 path is specialised for bottleneck width 128 / growth 32 (got %d / %d)",
                P.mid, P.d.growth);
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one chunk");
  // ---- conv1 (BN2 scale is folded into the bf16 weights by pack.cu; the epilogue adds the folded shift)
  if (which & 1) TCVN_TRY(launch_gemm(true, blk, rows, B.ctot, B.ld, pk + L.p_w1, kMid, L.kpad, L.kphys, pf(pk, L.p_a_scale),
                       pf(pk, L.p_a_shift), pf(pk, L.p_a_alpha), pf(pk, L.p_o_shift), pf(pk, L.p_o_alpha), mid, kMid, kMid,
                       1, B.Hp, B.Wp, st));
  // ---- conv2
  if (!(which & 2)) return TCVN_OK;
  return umma_conv2_fwd(mid, rows, pk + L.p_w2, pf(pk, L.p_b2), blk, B.ld, L.kphys, B.Hp, B.Wp, B.W, st);
}

// out[:, col0 : col0+32] = conv3x3(mid) + bias over the ringed layout; mid bf16 [rows][128] activated with a zero ring,
// w2 bf16 [9*32][128] (tap-major, K contiguous), out bf16 with pitch ldo
int umma_conv2_fwd(const void* mid, long long rows, const void* w2, const float* bias, void* out, int ldo, int col0, int Hp,
                   int Wp, int W, cudaStream_t st, float p_drop, unsigned long long seed, unsigned long long site,
                   double* stats, int* stat_slots) {
  if (rows >= (1ll << 31) - 4096) return fail(TCVN_ERR_UNSUPPORTED, "more than 2^31 rows in one chunk");
  const int tiles = (int)ceil_div_ll(rows, kC2Out);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (stat_slots) *stat_slots = grid;
  const int halo_rows_max = 288;
  bool& attr_done = device_flag(2);   // per device: the attribute belongs to the device's copy of the function
  if (!attr_done) {
    TCVN_CUDA(cudaFuncSetAttribute(umma_conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kC2SmemMax));
    attr_done = true;
  }
  Conv2Params c2;
  c2.m_total = rows; c2.Hp = Hp; c2.Wp = Wp;
  c2.nbox = ceil_div(kTileM + 2 * Wp, kBoxRows);
  c2.halo_rows = c2.nbox * kBoxRows;
  if (c2.halo_rows > halo_rows_max)
    return fail(TCVN_ERR_UNSUPPORTED, "feature map width %d needs a %d-row halo tile (max %d)", W, c2.halo_rows, halo_rows_max);
  {
    const size_t fixed = 1024 + kW2Bytes + 4864, per_stage = (size_t)2 * c2.halo_rows * 128;
    int st_n = (int)((kC2SmemMax - fixed) / per_stage);
    static const int st_cap = [] { const char* v = getenv("TCVN_C2_STAGES"); const int n = v ? atoi(v) : 0; return n >= 2 && n <= kC2StagesMax ? n : kC2StagesMax; }();
    c2.stages = st_n > st_cap ? st_cap : st_n;
    if (c2.stages < 2) return fail(TCVN_ERR_UNSUPPORTED, "conv2: a %d-row halo tile leaves no room for two stages", c2.halo_rows);
  }
  c2.bias = bias;
  c2.out = static_cast<bf16*>(out); c2.ldo = ldo; c2.col0 = col0; c2.num_tiles = tiles;
  c2.p_drop = p_drop; c2.seed = seed; c2.site = site; c2.seed_off = seed_offset_ptr(); c2.stats = stats;
  // Measured on B200: the 128B swizzle of a UMMA operand is a function of the absolute shared-memory address
  // (bits [4,7) ^= bits [7,10)), exactly as TMA wrote it, so a descriptor may start at ANY row of the haloed
  // tile with matrix-base-offset 0 (setting it to (addr >> 7) & 7 gives wrong results).
  CUtensorMap tmM, tmW2;
  TCVN_TRY(make_map(mid, rows, kMid, kMid, 64, kBoxRows, &tmM));
  TCVN_TRY(make_map(w2, 9 * kGrowth, kMid, kMid, 64, kC2N, &tmW2));
  const size_t smem2 = 1024 + kW2Bytes + (size_t)c2.stages * 2 * c2.halo_rows * 128 + 4864;
  umma_conv2_kernel<<<grid, kC2Threads, smem2, st>>>(tmM, tmW2, c2);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// transition 1x1 convolution on the pooled, activated map: N tiles of 128 output channels.  Columns beyond
// the real output width receive zeros (zero weight rows); in the next block's buffer those columns belong to
// dense layers that overwrite them before anything reads them, and the tensor map clips at the buffer width.
int umma_transition(const CnnPlan& P, const BlockPlan& B, const BlockPlan& Nx, const char* pk, const void* pool,
                    void* next_blk, long long rows, cudaStream_t st) {
  return launch_gemm(false, pool, rows, B.ctot, B.ctot, pk + B.p_tw16, B.tn_tiles * kMid, B.tkpad, B.ctot, nullptr,
                     nullptr, nullptr, pf(pk, B.p_tb16), pf(pk, B.p_ta16), next_blk, Nx.ctot, Nx.ld, B.tn_tiles, Nx.Hp,
                     Nx.Wp, st);
}

}  // namespace tcvn
