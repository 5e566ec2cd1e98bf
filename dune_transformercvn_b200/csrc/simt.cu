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
#include "stem.cuh"

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
  int a_ring_Hp, a_ring_Wp;  // > 0: activated A rows that fall on the zero ring read as 0 (conv zero padding)
  int accumulate;            // out += result
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

  // k-steps of all taps as one sequence; the operands of step i + 1 are loaded into registers while step i is computed
  // from shared memory (with the loads issued right before the barrier the small GEMMs of the sequence part - 8 to 24
  // steps - were a chain of exposed global-memory latencies: 22 us per launch whatever the size)
  float av[4];
  float4 bv;
  auto load = [&](int t, int k0) {
    const long long gm = m0 + ar + g.tap_off[t];
    bool row_ok = gm >= 0 && gm < g.m_total;
    if (row_ok && g.a_ring_Hp > 0) {
      const int rr = (int)(gm % (g.a_ring_Hp * g.a_ring_Wp));
      const int y = rr / g.a_ring_Wp, x = rr - y * g.a_ring_Wp;
      row_ok = !(y == 0 || y == g.a_ring_Hp - 1 || x == 0 || x == g.a_ring_Wp - 1);
    }
    const TA* arow = A + (row_ok ? gm : 0) * (long long)g.lda;
    const float* Wt = g.W + (size_t)t * g.K * g.N;
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
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bk < kBK && k0 + bk < g.K && n0 + bn < g.N)
      bv = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + bk) * g.N + n0 + bn));
  };
  int t = 0, k0 = 0;
  load(0, 0);
  while (t < g.taps) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) As[akq + i][ar] = av[i];
    if (bk < kBK) *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
    __syncthreads();
    k0 += kBK;
    if (k0 >= g.K) { k0 = 0; ++t; }
    if (t < g.taps) load(t, k0);
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
      else if (g.o_shift != nullptr) v += __ldg(g.o_shift + n);
      if (ring) v = 0.f;
      TO* dst = out + m * (long long)g.ldo + g.out_col0 + n;
      if (g.accumulate) v += to_f32<TO>(*dst);
      *dst = from_f32<TO>(v);
    }
  }
}

int launch_simt_gemm(const GemmArgs& a, cudaStream_t stream) {
  TCVN_CHECK_ARG(a.N % 4 == 0 && a.taps >= 1 && a.taps <= 9, "simt_gemm: bad arguments");
  if (a.m_total <= 0) return TCVN_OK;
  GemmDev g;
  g.A = a.A; g.lda = a.lda; g.m_total = a.m_total; g.K = a.K; g.taps = a.taps;
  for (int i = 0; i < 9; ++i) g.tap_off[i] = a.tap_off[i];
  g.W = a.W; g.N = a.N;
  g.a_scale = a.a_scale; g.a_shift = a.a_shift; g.a_alpha = a.a_alpha;
  g.o_scale = a.o_scale; g.o_shift = a.o_shift; g.o_alpha = a.o_alpha;
  g.out = a.out; g.ldo = a.ldo; g.out_col0 = a.out_col0; g.ring_Hp = a.ring_Hp; g.ring_Wp = a.ring_Wp;
  g.a_ring_Hp = a.a_ring_Hp; g.a_ring_Wp = a.a_ring_Wp; g.accumulate = a.accumulate ? 1 : 0;
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
// Stem, fused and hit-driven: 7x7 stride-2 convolution on NCHW fp32 pixels + bias + BN0 + PReLU0 +
// AvgPool2d(3, 2), written straight into channels [0, 64) of block 0's ringed buffer
// (dense_net.py:111-122).  The 64 x 200 x 140 pre-pool map never reaches memory.
//
// The pixel maps are ~99 % zeros, so the convolution is evaluated from the hits: a persistent CTA keeps
// the 147 x 64 filter bank in shared memory and loops over 8 x 8 tiles of pooled pixels
// (17 x 17 conv outputs <- 39 x 39 x 3 input window).  Per tile it (1) compacts the non-zero pixels
// of the window into a list in a fixed order (ballot + prefix sums: deterministic), (2) lets every hit
// add v * w[tap] to the <= 4 x 4 conv outputs it reaches - thread (cy mod 4, cx mod 2, channel) owns
// its accumulators, so there are no atomics and no barriers between hits, and the summation order is
// fixed, (3) applies bias/BN/PReLU in place and (4) average-pools 3x3/2 into the output rows.  For a
// dense input this degrades gracefully to the full convolution (the list holds the whole window).
// ------------------------------------------------------------------------------------------------
template <typename TO, int C0>
__device__ __forceinline__ void stem_pool(const float* acc, const int* touched, float sc, float sh, float al, int ch, int t, int n,
                                          int py0, int px0, TO* __restrict__ blk, int ldo, int Hb, int Wb) {
  // ---- (3)+(4) bias + BN0 + PReLU0, AvgPool2d(3, 2) -> ringed block buffer.  A conv output no hit reached is
  // the per-channel constant PReLU(shift), so a pooling window of nine such outputs is a per-channel constant
  // too (evaluated with the same additions and division as the general case: the shortcut is bit-identical).
  // Pixels whose window was reached are listed and evaluated 8 at a time (64 channels x 8 pixels = 512 threads);
  // all other pixels get the constant vector with 16-byte stores.
  __shared__ int pool_list[kStemTP * kStemTP];
  __shared__ unsigned char pool_kind[kStemTP * kStemTP];  // 0 constant, 1 reached (listed), 2 outside the map
  __shared__ int pool_count;
  __shared__ __align__(16) TO pool_const[C0];
  const float c_act = prelu(fmaf(0.f, sc, sh), al);
  float c_pool = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) c_pool += c_act;
  c_pool = pool_avg9<TO>(c_pool);
  if (t < C0) pool_const[t] = from_f32<TO>(c_pool);
  if (t < 32) {  // one warp classifies the 64 pooled pixels (2 per lane) and compacts the reached ones in order
    int cnt = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int p = half * 32 + t;
      const int pyl = p / kStemTP, pxl = p % kStemTP;
      const int base = (2 * pyl) * kStemTC + 2 * pxl;
      int any = 0;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) any |= touched[base + dy * kStemTC + dx];
      const bool inside = py0 + pyl < Hb && px0 + pxl < Wb;
      const unsigned b = __ballot_sync(0xffffffffu, any && inside);
      if (any && inside) pool_list[cnt + __popc(b & ((1u << t) - 1u))] = p;
      pool_kind[p] = inside ? (any ? 1 : 0) : 2;
      cnt += __popc(b);
    }
    if (t == 0) pool_count = cnt;
  }
  __syncthreads();
  // constant pixels: 64 pixels x 8 sixteen-byte chunks
  {
    constexpr int VEC = 16 / sizeof(TO);           // channels per 16-byte store
    constexpr int CHUNKS = C0 / VEC;               // stores per pixel
    for (int i = t; i < kStemTP * kStemTP * CHUNKS; i += blockDim.x) {
      const int p = i / CHUNKS, q = i % CHUNKS;
      const int pyl = p / kStemTP, pxl = p % kStemTP;
      const int py = py0 + pyl, px = px0 + pxl;
      if (pool_kind[p] != 0) continue;
      const size_t row = (size_t)n * (Hb + 2) * (Wb + 2) + (size_t)(py + 1) * (Wb + 2) + (px + 1);
      *reinterpret_cast<uint4*>(blk + row * ldo + q * VEC) = *reinterpret_cast<const uint4*>(pool_const + q * VEC);
    }
  }
  // reached pixels: full evaluation, one pixel per group of C0 threads
  const int npool = pool_count;
  for (int idx = t / C0; idx < npool; idx += blockDim.x / C0) {
    const int p = pool_list[idx];
    const int pyl = p / kStemTP, pxl = p % kStemTP;
    const int base = (2 * pyl) * kStemTC + 2 * pxl;
    float s2 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) s2 += prelu(fmaf(acc[(base + dy * kStemTC + dx) * C0 + ch], sc, sh), al);
    const size_t row = (size_t)n * (Hb + 2) * (Wb + 2) + (size_t)(py0 + pyl + 1) * (Wb + 2) + (px0 + pxl + 1);
    blk[row * ldo + ch] = from_f32<TO>(pool_avg9<TO>(s2));
  }
}

// phases (2)-(4) back to back (dense-window and un-binned COO kernels)
template <typename TO, int C0, bool WGLOBAL = false>
__device__ __forceinline__ void stem_scatter_pool(const float* wsm, float* acc, const float4* hits, int nhits,
                                                  int* touched, int cin, float sc, float sh, float al, int ch, int own_py,
                                                  int own_px, int t, int n, int py0, int px0, TO* __restrict__ blk,
                                                  int ldo, int Hb, int Wb) {
  stem_scatter<C0, WGLOBAL>(wsm, acc, hits, nhits, touched, cin, ch, own_py, own_px);
  __syncthreads();
  stem_pool<TO, C0>(acc, touched, sc, sh, al, ch, t, n, py0, px0, blk, ldo, Hb, Wb);
}

template <typename TO, int C0>
__global__ void __launch_bounds__(kStemThreads) stem_fused_kernel(const float* __restrict__ pixels, int n_images, int cin,
                                                                  int H, int W, const float* __restrict__ w0,
                                                                  const float* __restrict__ s_scale,
                                                                  const float* __restrict__ s_shift,
                                                                  const float* __restrict__ s_alpha,
                                                                  TO* __restrict__ blk, int ldo, int Hb, int Wb) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;                                  // [cin*49][C0]
  float* acc = wsm + cin * 49 * C0;                   // [289][C0]
  float4* hits = reinterpret_cast<float4*>(acc + kStemTC * kStemTC * C0);  // [39*39] (packed yx, v0, v1, v2)
  __shared__ int row_count[kStemRows];
  __shared__ int row_start[kStemRows + 1];
  __shared__ int touched[kStemTC * kStemTC];  // conv outputs reached by at least one hit of this tile
  for (int i = threadIdx.x; i < cin * 49 * C0; i += blockDim.x) wsm[i] = __ldg(w0 + i);
  const int tiles_x = (Wb + kStemTP - 1) / kStemTP, tiles_y = (Hb + kStemTP - 1) / kStemTP;
  const int per_image = tiles_x * tiles_y;
  const long long total = (long long)n_images * per_image;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int ch = t & (C0 - 1);
  const int own_px = (t >> 6) & 1, own_py = t >> 7;   // owns conv outputs with cx % 2 == own_px, cy % 4 == own_py
  const float sc = __ldg(s_scale + ch), sh = __ldg(s_shift + ch), al = __ldg(s_alpha + ch);
  const size_t plane = (size_t)H * W;
  // window values of the tile being compacted: [row slot][column slot][channel]; warp w scans window rows
  // w, w+16, w+32, lane covers columns lane and lane+32.  Loaded one tile AHEAD (pure loads, no votes in
  // between) so the global-memory latency hides behind the previous tile's scatter / pooling.
  float v[3][2][3];
  auto load_window = [&](long long tile_id) {
    const int n = (int)(tile_id / per_image);
    const int rem = (int)(tile_id - (long long)n * per_image);
    const int iy0 = 4 * (rem / tiles_x) * kStemTP - 3, ix0 = 4 * (rem % tiles_x) * kStemTP - 3;
    const float* img = pixels + (size_t)n * cin * plane;
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs, y = iy0 + yy;
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        const int xx = lane + 32 * cs, x = ix0 + xx;
        const bool ok = yy < kStemIn && xx < kStemIn && y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[rs][cs][c] = (ok && c < cin) ? __ldg(img + c * plane + (size_t)y * W + x) : 0.f;
      }
    }
  };
  if ((long long)blockIdx.x < total) load_window(blockIdx.x);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = (int)(tile / per_image);
    const int rem = (int)(tile - (long long)n * per_image);
    const int py0 = (rem / tiles_x) * kStemTP, px0 = (rem % tiles_x) * kStemTP;
    __syncthreads();  // previous tile fully consumed
    // ---- (1) compact the non-zero pixels of the window, rows in order, columns in order
    unsigned m0[3], m1[3];
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs;
      m0[rs] = __ballot_sync(0xffffffffu, v[rs][0][0] != 0.f || v[rs][0][1] != 0.f || v[rs][0][2] != 0.f);
      m1[rs] = __ballot_sync(0xffffffffu, v[rs][1][0] != 0.f || v[rs][1][1] != 0.f || v[rs][1][2] != 0.f);
      if (lane == 0 && yy < kStemIn) row_count[yy] = __popc(m0[rs]) + __popc(m1[rs]);
    }
    // zero the accumulators while the counts settle
    {
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
#pragma unroll
    for (int rs = 0; rs < 3; ++rs) {
      const int yy = warp + 16 * rs;
      if (yy < kStemIn) {
        const int base = row_start[yy];
        const unsigned below = (1u << lane) - 1u;
        if (m0[rs] >> lane & 1u)
          hits[base + __popc(m0[rs] & below)] =
              make_float4(__int_as_float(yy * 64 + lane), v[rs][0][0], v[rs][0][1], v[rs][0][2]);
        if (m1[rs] >> lane & 1u)
          hits[base + __popc(m0[rs]) + __popc(m1[rs] & below)] =
              make_float4(__int_as_float(yy * 64 + lane + 32), v[rs][1][0], v[rs][1][1], v[rs][1][2]);
      }
    }
    if (tile + gridDim.x < total) load_window(tile + gridDim.x);  // in flight during the rest of this tile
    __syncthreads();
    stem_scatter_pool<TO, C0>(wsm, acc, hits, row_start[kStemRows], touched, cin, sc, sh, al, ch, own_py, own_px, t, n, py0,
                              px0, blk, ldo, Hb, Wb);
  }
}

// ------------------------------------------------------------------------------------------------
// Stem straight from the Minkowski-format hit list: the dense 3 x 400 x 280 map is never built.
// Fuses sparse_to_dense + "/255" (neutrino_full_dense_trainer.py:15-24,59-60) into the stem: each tile CTA
// walks its image's slice of the (image-sorted) hit list in order, keeps the hits that fall into its input
// window (ballot compaction: deterministic order) and hands them to the same scatter / pool code.
// image_offsets[i] = first hit of image i (built by hit_offsets_kernel), so no binary search per tile.
// ------------------------------------------------------------------------------------------------
__global__ void hit_offsets_kernel(const int32_t* __restrict__ coords, long long nnz, int n_images,
                                   long long* __restrict__ offsets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_images) return;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(coords + 3 * mid) < i) lo = mid + 1; else hi = mid;
  }
  offsets[i] = lo;
}

template <typename TO, int C0, typename V>
__global__ void __launch_bounds__(kStemThreads, 2) stem_coo_kernel(const int32_t* __restrict__ coords,
                                                                const V* __restrict__ values,
                                                                const long long* __restrict__ image_offsets, int image0,
                                                                float divisor, int n_images, int cin, int H, int W,
                                                                const float* __restrict__ w0,
                                                                const float* __restrict__ s_scale,
                                                                const float* __restrict__ s_shift,
                                                                const float* __restrict__ s_alpha, TO* __restrict__ blk,
                                                                int ldo, int Hb, int Wb) {
  // the kernel is issue/latency-bound (ncu: 41 % issue slots busy at 25 % occupancy with one 136 KB CTA per SM): the
  // filter bank is read from global memory through the read-only path instead of a shared copy, which brings the
  // CTA down to 98 KB and two CTAs onto every SM
  extern __shared__ __align__(16) float smem[];
  float* acc = smem;                                  // [289][C0]
  float4* hits = reinterpret_cast<float4*>(acc + kStemTC * kStemTC * C0);  // [39*39] (packed yx, v0, v1, v2)
  __shared__ int wcount[kStemThreads / 32];
  __shared__ int touched[kStemTC * kStemTC];
  const int tiles_x = (Wb + kStemTP - 1) / kStemTP, tiles_y = (Hb + kStemTP - 1) / kStemTP;
  const int per_image = tiles_x * tiles_y;
  const long long total = (long long)n_images * per_image;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int ch = t & (C0 - 1);
  const int own_px = (t >> 6) & 1, own_py = t >> 7;
  const float sc = __ldg(s_scale + ch), sh = __ldg(s_shift + ch), al = __ldg(s_alpha + ch);
  constexpr int kCap = kStemIn * kStemIn;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = (int)tile / per_image;   // launch_stem_coo keeps the tile count below 2^31
    const int rem = (int)tile - n * per_image;
    const int py0 = (rem / tiles_x) * kStemTP, px0 = (rem % tiles_x) * kStemTP;
    const int iy0 = 4 * py0 - 3, ix0 = 4 * px0 - 3;
    const long long lo = __ldg(image_offsets + image0 + n), hi = __ldg(image_offsets + image0 + n + 1);
    __syncthreads();  // previous tile fully consumed
    {
      float4* a4 = reinterpret_cast<float4*>(acc);
      for (int i = t; i < kStemTC * kStemTC * C0 / 4; i += blockDim.x) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < kStemTC * kStemTC) touched[t] = 0;
    }
    int base = 0;
    for (long long h0 = lo; h0 < hi; h0 += kStemThreads) {
      const long long h = h0 + t;
      bool in = false;
      int yy = 0, xx = 0;
      if (h < hi) {
        const int y = __ldg(coords + 3 * h + 1), x = __ldg(coords + 3 * h + 2);
        yy = y - iy0; xx = x - ix0;
        in = y >= 0 && y < H && x >= 0 && x < W && yy >= 0 && yy < kStemIn && xx >= 0 && xx < kStemIn;
      }
      const unsigned b = __ballot_sync(0xffffffffu, in);
      if (lane == 0) wcount[warp] = __popc(b);
      __syncthreads();
      int before = 0, all = 0;
#pragma unroll
      for (int w = 0; w < kStemThreads / 32; ++w) {
        const int c = wcount[w];
        before += w < warp ? c : 0;
        all += c;
      }
      if (in) {
        const int slot = base + before + __popc(b & ((1u << lane) - 1u));
        float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < cin) {
            float val = static_cast<float>(values[h * cin + c]);
            if (divisor != 0.f) val = __fdiv_rn(val, divisor);  // same bits as the reference's v / 255.0
            v[c] = val;
          }
        if (slot < kCap) hits[slot] = make_float4(__int_as_float(yy * 64 + xx), v[0], v[1], v[2]);
      }
      base += all;
      __syncthreads();
    }
    __syncthreads();
    stem_scatter_pool<TO, C0, true>(w0, acc, hits, base < kCap ? base : kCap, touched, cin, sc, sh, al, ch, own_py, own_px, t, n,
                                    py0, px0, blk, ldo, Hb, Wb);
  }
}

int launch_hit_offsets(const int32_t* coords, long long nnz, int n_images, long long* offsets, cudaStream_t stream) {
  hit_offsets_kernel<<<ceil_div(n_images + 1, 128), 128, 0, stream>>>(coords, nnz, n_images, offsets);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int launch_stem_coo(const int32_t* coords, const void* values, bool values_u8, const long long* image_offsets, int image0,
                    float divisor, int n, int cin, int H, int W, const float* w0, const float* s_scale,
                    const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb, int Wb, bool f32,
                    cudaStream_t stream) {
  if (c0 != 64) return fail(TCVN_ERR_UNSUPPORTED, "stem: init_features %d (kernel is specialised for 64)", c0);
  if (n == 0) return TCVN_OK;
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  if ((Hs - 3) / 2 + 1 != Hb || (Ws - 3) / 2 + 1 != Wb) return fail(TCVN_ERR_ARG, "stem: geometry mismatch");
  if (cin > 3) return fail(TCVN_ERR_UNSUPPORTED, "stem: %d input channels (kernel handles up to 3)", cin);
  const size_t smem = (size_t)kStemTC * kStemTC * c0 * sizeof(float) + (size_t)kStemIn * kStemIn * sizeof(float4);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long tiles = (long long)n * ((Wb + kStemTP - 1) / kStemTP) * ((Hb + kStemTP - 1) / kStemTP);
  if (tiles >= (1ll << 31)) return fail(TCVN_ERR_UNSUPPORTED, "stem: too many tiles in one chunk");
  const int grid = (int)(tiles < 2ll * sms ? tiles : 2ll * sms);   // two resident CTAs per SM
#define TCVN_STEM_COO(TO, V)                                                                                          \
  do {                                                                                                                \
    TCVN_CUDA(cudaFuncSetAttribute(stem_coo_kernel<TO, 64, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    TCVN_CUDA(cudaFuncSetAttribute(stem_coo_kernel<TO, 64, V>, cudaFuncAttributePreferredSharedMemoryCarveout,          \
                                   (int)cudaSharedmemCarveoutMaxShared));                                               \
    stem_coo_kernel<TO, 64, V><<<grid, kStemThreads, smem, stream>>>(coords, static_cast<const V*>(values), image_offsets, \
                                                                     image0, divisor, n, cin, H, W, w0, s_scale, s_shift, \
                                                                     s_alpha, static_cast<TO*>(blk), ldo, Hb, Wb);     \
  } while (0)
  if (f32 && !values_u8) TCVN_STEM_COO(float, float);
  else if (f32 && values_u8) TCVN_STEM_COO(float, uint8_t);
  else if (!f32 && !values_u8) TCVN_STEM_COO(__nv_bfloat16, float);
  else TCVN_STEM_COO(__nv_bfloat16, uint8_t);
#undef TCVN_STEM_COO
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int launch_stem(const float* pixels, int n, int cin, int H, int W, const float* w0, const float* s_scale,
                const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb, int Wb, bool f32,
                cudaStream_t stream) {
  if (c0 != 64) return fail(TCVN_ERR_UNSUPPORTED, "stem: init_features %d (kernel is specialised for 64)", c0);
  if (n == 0) return TCVN_OK;
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  if ((Hs - 3) / 2 + 1 != Hb || (Ws - 3) / 2 + 1 != Wb) return fail(TCVN_ERR_ARG, "stem: geometry mismatch");
  if (cin > 3) return fail(TCVN_ERR_UNSUPPORTED, "stem: %d input channels (kernel handles up to 3)", cin);
  const size_t smem = ((size_t)cin * 49 * c0 + (size_t)kStemTC * kStemTC * c0) * sizeof(float) +
                      (size_t)kStemIn * kStemIn * sizeof(float4);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long tiles = (long long)n * ((Wb + kStemTP - 1) / kStemTP) * ((Hb + kStemTP - 1) / kStemTP);
  const int grid = (int)(tiles < (long long)sms ? tiles : (long long)sms);
  if (f32) {
    TCVN_CUDA(cudaFuncSetAttribute(stem_fused_kernel<float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_fused_kernel<float, 64><<<grid, kStemThreads, smem, stream>>>(pixels, n, cin, H, W, w0, s_scale, s_shift, s_alpha,
                                                                        static_cast<float*>(blk), ldo, Hb, Wb);
  } else {
    TCVN_CUDA(cudaFuncSetAttribute(stem_fused_kernel<__nv_bfloat16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    stem_fused_kernel<__nv_bfloat16, 64><<<grid, kStemThreads, smem, stream>>>(
        pixels, n, cin, H, W, w0, s_scale, s_shift, s_alpha, static_cast<__nv_bfloat16*>(blk), ldo, Hb, Wb);
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// Transition front half.  AvgPool2d(2,2) is linear and the 1x1 convolution is per-pixel, so
// pool(conv(a)) == conv(pool(a)) (bias included): pooling the activated map first makes the GEMM 4x smaller.
// one thread = one output pixel x 8 (bf16) / 4 (fp32) channels: 16-byte loads and stores
template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[Vec16<T>::N]);
template <>
__device__ __forceinline__ void load_vec<float>(const float* p, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[Vec16<T>::N]);
template <>
__device__ __forceinline__ void store_vec<float>(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Persistent: a thread keeps ONE group of V channels (its 3 x V BN/PReLU constants stay in registers) and walks over
// output pixels; with one (pixel, channel group) per thread the 24 scalar constant loads per thread - lanes 32 bytes
// apart, 8 wavefronts each - were ten times the LSU work of the four 16-byte data loads (59 % of the HBM bandwidth).
// blockDim.x = cv * rows with cv = c / V channel groups; lanes run over channel groups, so a pixel is read contiguously.
template <typename T>
__global__ void __launch_bounds__(256) act_pool2_kernel(const T* __restrict__ blk, int H, int W, int ld, int c,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        const float* __restrict__ alpha, T* __restrict__ out, int H2, int W2,
                                                        long long n_pixels, int cv) {
  constexpr int V = Vec16<T>::N;
  const int v = threadIdx.x % cv, ty = threadIdx.x / cv, rows = blockDim.x / cv;
  const int ch = v * V;
  float sc[V], sh[V], al[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { sc[i] = __ldg(scale + ch + i); sh[i] = __ldg(shift + ch + i); al[i] = __ldg(alpha + ch + i); }
  const int Wp = W + 2;
  const long long per_image = (long long)H2 * W2;
  for (long long px = (long long)blockIdx.x * rows + ty; px < n_pixels; px += (long long)gridDim.x * rows) {
    const int n = (int)(px / per_image);
    const int r = (int)(px - (long long)n * per_image);
    const int y = r / W2, x = r - y * W2;
    const size_t row0 = (size_t)n * (H + 2) * Wp + (size_t)(2 * y + 1) * Wp + (2 * x + 1);
    float f[4][V];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) load_vec<T>(blk + (row0 + (size_t)dy * Wp + dx) * ld + ch, f[dy * 2 + dx]);
    float s[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += prelu(fmaf(f[q][i], sc[i], sh[i]), al[i]);
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] *= 0.25f;
    const size_t orow = (size_t)n * (H2 + 2) * (W2 + 2) + (size_t)(y + 1) * (W2 + 2) + (x + 1);
    store_vec<T>(out + orow * c + ch, s);
  }
}

int launch_act_pool2(const void* blk, int n, int H, int W, int ld, int c, const float* scale, const float* shift,
                     const float* alpha, void* out, int H2, int W2, bool f32, cudaStream_t stream) {
  const int V = f32 ? 4 : 8;
  if (c % V || ld % V) return fail(TCVN_ERR_UNSUPPORTED, "act_pool2: channel count %d / pitch %d not a multiple of %d", c, ld, V);
  const int cv = c / V;
  if (cv > 256) return fail(TCVN_ERR_UNSUPPORTED, "act_pool2: %d channels (at most %d)", c, 256 * V);
  const long long n_pixels = (long long)n * H2 * W2;
  if (n_pixels == 0) return TCVN_OK;
  const int rows = 256 / cv;
  const int threads = rows * cv;
  long long blocks = ceil_div_ll(n_pixels, (long long)rows * 4);   // >= 4 pixels per thread
  const long long cap = 148ll * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (f32) act_pool2_kernel<float><<<(unsigned)blocks, threads, 0, stream>>>(static_cast<const float*>(blk), H, W, ld, c, scale,
                                                                             shift, alpha, static_cast<float*>(out), H2, W2,
                                                                             n_pixels, cv);
  else act_pool2_kernel<__nv_bfloat16><<<(unsigned)blocks, threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(blk), H, W, ld, c, scale, shift, alpha, static_cast<__nv_bfloat16*>(out), H2, W2, n_pixels,
      cv);
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
