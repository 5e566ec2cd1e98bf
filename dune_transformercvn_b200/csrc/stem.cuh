// Shared device code of the hit-driven stem (simt.cu: inference kernels; train_stem.cu: training kernels).
// conv 7x7 stride 2 pad 3 evaluated from the non-zero pixels of a tile's input window; see simt.cu for the description.
#pragma once
#include "common.cuh"

namespace tcvn {

constexpr int kStemTP = 8;                   // pooled tile edge
constexpr int kStemTC = 2 * kStemTP + 1;     // 17 conv outputs per edge
constexpr int kStemIn = 2 * kStemTC + 5;     // 39 input pixels per edge
constexpr int kStemThreads = 512;            // (cy mod 4) x (cx mod 2) x 64 channels
constexpr int kStemRows = kStemIn;           // window rows (pixels, all channels together)

// AvgPool2d(3, 2) divides the window sum by 9.  fp32 path: the reference's division.  bf16 path: the product with 1/9
// (at most one fp32 ulp away, far below the bf16 rounding that follows; a division is ~10 instructions per value).
template <typename TO> __device__ __forceinline__ float pool_avg9(float s) { return s / 9.0f; }
template <> __device__ __forceinline__ float pool_avg9<__nv_bfloat16>(float s) { return s * (1.0f / 9.0f); }
template <typename TO> __device__ __forceinline__ float2 pool_avg9(float2 s) {
  return make_float2(pool_avg9<TO>(s.x), pool_avg9<TO>(s.y));
}

// Phases (2)-(4) of the stem for one tile, shared by the dense-window and the COO-direct kernels: scatter the
// compacted hits, then bias+BN0+PReLU0 and AvgPool2d(3, 2) into the ringed block buffer.
template <int C0, bool WGLOBAL>
__device__ __forceinline__ void stem_scatter(const float* wsm, float* acc, const float4* hits, int nhits, int* touched, int cin,
                                             int ch, int own_py, int own_px) {
  // ---- (2) scatter every hit into the conv outputs it reaches (input yy = 2*cy + ky)
  for (int h = 0; h < nhits; ++h) {
    const float4 hit = hits[h];
    const int code = __float_as_int(hit.x);
    const int yy = code >> 6, xx = code & 63;
    // the unique cy in [cy_lo, cy_lo + 4) with cy % 4 == own_py, cy_lo = max(0, ceil((yy - 6) / 2))
    const int cy_lo = yy > 6 ? (yy - 5) >> 1 : 0;
    const int cy = cy_lo + ((own_py - cy_lo) & 3);
    const int ky = yy - 2 * cy;
    if (cy > kStemTC - 1 || ky < 0) continue;   // ky <= 6 by construction
    const int cx_lo = xx > 6 ? (xx - 5) >> 1 : 0;
    const int cx_a = cx_lo + ((own_px - cx_lo) & 1);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int cx = cx_a + 2 * k;
      const int kx = xx - 2 * cx;
      if (cx <= kStemTC - 1 && kx >= 0) {
        const float* wp = wsm + (ky * 7 + kx) * C0 + ch;
        float a = acc[(cy * kStemTC + cx) * C0 + ch];
        // WGLOBAL: the 37.6 KB filter bank stays in global memory (L1-resident, read-only path) so that two CTAs fit an SM
        a = fmaf(hit.y, WGLOBAL ? __ldg(wp) : wp[0], a);
        if (cin > 1) a = fmaf(hit.z, WGLOBAL ? __ldg(wp + 49 * C0) : wp[49 * C0], a);
        if (cin > 2) a = fmaf(hit.w, WGLOBAL ? __ldg(wp + 2 * 49 * C0) : wp[2 * 49 * C0], a);
        acc[(cy * kStemTC + cx) * C0 + ch] = a;
        if (ch == 0) touched[cy * kStemTC + cx] = 1;
      }
    }
  }
}

}  // namespace tcvn
