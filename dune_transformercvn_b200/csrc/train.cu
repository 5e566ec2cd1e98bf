// Training primitives (fp32, CUDA cores): the building blocks of the train-mode forward and the hand-written
// backward of the hot path.  Everything works on row matrices: a feature map is the ringed channels-last
// matrix [N*(H+2)*(W+2), C] of the inference path (ring rows carry no data and no gradient), token / image
// level tensors are plain [rows, C] matrices (ring_hp == 0).
//
// Reference semantics being reproduced (torch defaults):
//   BatchNorm (dense_net.py:19,30,85,119,147,159; prong_feature_embedding.py:17; encoder.py:14) in train mode:
//     normalise with the batch mean and the BIASED batch variance, update running_mean / running_var (the
//     latter with the UNBIASED variance) with momentum 0.1;
//   PReLU with one slope per channel; backward d(alpha) = sum(dy * min(x, 0)).
#include "kernels.h"
#include "umma.h"

namespace tcvn {

__device__ __forceinline__ bool is_ring(long long m, int Hp, int Wp) {
  if (Hp <= 0) return false;
  // every launcher bounds the row count below 2^31: 32-bit division
  const unsigned rr = (unsigned)m % (unsigned)(Hp * Wp);
  const unsigned y = rr / (unsigned)Wp, x = rr - y * (unsigned)Wp;
  return y == 0 || y == (unsigned)(Hp - 1) || x == 0 || x == (unsigned)(Wp - 1);
}

// ------------------------------------------------------------------------------------------------
// weight gradient of the (shifted) GEMM:  dW[t][k][n] += sum_m act(A[m + off_t, k]) * G[m, n]
// ------------------------------------------------------------------------------------------------
struct WgradDev {
  const void* A; int lda; long long m_total; int K; int taps; int tap_off[9];
  const float *a_scale, *a_shift, *a_alpha;
  int a_ring_Hp, a_ring_Wp;
  const void* G; int ldg; int g_col0; int N;
  int g_ring_Hp, g_ring_Wp;
  float* dW;  // [taps][K][N], accumulated (atomics when several row slabs share it)
  float* parts;  // non-null: slab z stores its partial result to parts[z][taps][K][N] instead (added later in a fixed order)
  int rows_per_slab;
};

constexpr int kWgK = 64, kWgN = 32, kWgR = 32;

template <typename TA, typename TG>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradDev g) {
  const TA* gA = static_cast<const TA*>(g.A);
  const TG* gG = static_cast<const TG*>(g.G);
  __shared__ float As[kWgR][kWgK + 1];
  __shared__ float Gs[kWgR][kWgN + 1];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * kWgK;
  const int n0 = (blockIdx.y % ((g.N + kWgN - 1) / kWgN)) * kWgN;
  const int tap = blockIdx.y / ((g.N + kWgN - 1) / kWgN);
  const long long r_begin = (long long)blockIdx.z * g.rows_per_slab;
  const long long r_end = min(g.m_total, r_begin + g.rows_per_slab);
  const int tk = (tid >> 4) * 4, tn = (tid & 15) * 2;  // 16 k-groups x 16 n-groups
  float acc[4][2] = {};
  const bool transform = g.a_scale != nullptr;
  for (long long r0 = r_begin; r0 < r_end; r0 += kWgR) {
    // A tile: 32 rows x 64 k, 8 elements per thread (row = tid / 8, k = (tid % 8) * 8 ..)
    {
      const int r = tid >> 3, kq = (tid & 7) * 8;
      const long long m = r0 + r;
      const long long gm = m + g.tap_off[tap];
      const bool ok = m < r_end && gm >= 0 && gm < g.m_total && !is_ring(gm, g.a_ring_Hp, g.a_ring_Wp);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + kq + i;
        float v = 0.f;
        if (ok && k < g.K) {
          v = to_f32<TA>(gA[gm * (long long)g.lda + k]);
          if (transform) v = prelu(fmaf(v, __ldg(g.a_scale + k), __ldg(g.a_shift + k)), __ldg(g.a_alpha + k));
        }
        As[r][kq + i] = v;
      }
    }
    {
      const int r = tid >> 3, nq = (tid & 7) * 4;
      const long long m = r0 + r;
      const bool ok = m < r_end && !is_ring(m, g.g_ring_Hp, g.g_ring_Wp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + nq + i;
        Gs[r][nq + i] = (ok && n < g.N) ? to_f32<TG>(gG[m * (long long)g.ldg + g.g_col0 + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kWgR; ++r) {
      const float g0 = Gs[r][tn], g1 = Gs[r][tn + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = As[r][tk + i];
        acc[i][0] = fmaf(a, g0, acc[i][0]);
        acc[i][1] = fmaf(a, g1, acc[i][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + tk + i, n = n0 + tn + j;
      if (k < g.K && n < g.N) {
        const size_t o = ((size_t)tap * g.K + k) * g.N + n;
        if (g.parts) g.parts[(size_t)blockIdx.z * g.taps * g.K * g.N + o] = acc[i][j];
        else atomicAdd(g.dW + o, acc[i][j]);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// per-column sums over the interior rows:  out[j][c] += sum_m f_j(...)
//   MODE 0  X                       -> sum x, sum x^2                    (BatchNorm batch statistics)
//   MODE 1  dA, X, fold             -> sum g, sum g*xhat, sum dA*min(y,0) (BN + PReLU backward reductions)
//   MODE 2  X                       -> sum x                             (bias gradients)
// fold = [scale | shift | alpha | mean | rstd] (5 x C), y = scale*x + shift, g = dA * (y >= 0 ? 1 : alpha)
// ------------------------------------------------------------------------------------------------
struct ColDev {
  const void* X; int ldx; int xcol0;
  const void* D; int ldd; int dcol0;
  const float* fold; int fold_stride; int C;
  long long m_total; int Hp, Wp; int rows_per_slab;
  // slab (blockIdx.y) s writes its partial sum j of column c to dst[s * slot_stride + j * sum_stride + c] with a plain
  // store: no atomics, so a fixed-order reduction over the slabs (parts_reduce_kernel, or fused into the consumer) makes
  // every statistic bit-reproducible
  double* dst; long long slot_stride; int sum_stride;
};

struct BnBwdDev {
  const void* D; int ldd; int dcol0;
  const void* X; int ldx; int xcol0;
  const float* fold; int fold_stride; const double* sums; int C; double count;
  void* dX; int lddx; int dxcol0; int accumulate;
  long long m_total; int Hp, Wp;
};

template <int MODE, typename TX, typename TD>
__global__ void __launch_bounds__(256) colsum_kernel(const ColDev p) {
  constexpr int NS = MODE == 0 ? 2 : (MODE == 1 ? 3 : 1);
  __shared__ double red[8][32][NS];
  const TX* X = static_cast<const TX*>(p.X);
  const TD* D = static_cast<const TD*>(p.D);
  const int lane_c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane_c;
  const long long r_begin = (long long)blockIdx.y * p.rows_per_slab;
  const long long r_end = min(p.m_total, r_begin + p.rows_per_slab);
  double acc[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) acc[j] = 0.0;
  if (c < p.C) {
    float sc = 0.f, sh = 0.f, al = 0.f, mean = 0.f, rstd = 0.f;
    if (MODE == 1) {
      const int fs = p.fold_stride;
      sc = p.fold[c]; sh = p.fold[fs + c]; al = p.fold[2 * fs + c]; mean = p.fold[3 * fs + c]; rstd = p.fold[4 * fs + c];
    }
    float part[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) part[j] = 0.f;
    int cnt = 0;
    for (long long m = r_begin + rl; m < r_end; m += 8) {
      if (!is_ring(m, p.Hp, p.Wp)) {
        if (MODE == 0) {
          const float x = to_f32<TX>(X[m * (long long)p.ldx + p.xcol0 + c]);
          part[0] += x;
          part[1] = fmaf(x, x, part[1]);
        } else if (MODE == 1) {
          const float x = to_f32<TX>(X[m * (long long)p.ldx + p.xcol0 + c]);
          const float d = to_f32<TD>(D[m * (long long)p.ldd + p.dcol0 + c]);
          const float y = fmaf(x, sc, sh);
          const float g = y >= 0.f ? d : d * al;
          part[0] += g;
          part[1] = fmaf(g, (x - mean) * rstd, part[1]);
          part[2] = fmaf(d, fminf(y, 0.f), part[2]);
        } else {
          part[0] += to_f32<TX>(X[m * (long long)p.ldx + p.xcol0 + c]);
        }
      }
      if (++cnt == 64) {  // flush the fp32 partials into doubles every 64 rows
#pragma unroll
        for (int j = 0; j < NS; ++j) { acc[j] += (double)part[j]; part[j] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] += (double)part[j];
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) red[rl][lane_c][j] = acc[j];
  __syncthreads();
  if (rl == 0 && c < p.C) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double s = 0.0;
      for (int r = 0; r < 8; ++r) s += red[r][lane_c][j];
      p.dst[(size_t)blockIdx.y * p.slot_stride + (size_t)j * p.sum_stride + c] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// vectorised variants (8 channels = one 16-byte bf16 / two 16-byte fp32 loads per thread, several rows in flight):
// used whenever C, the pitches and the column offsets are multiples of 8.  A block is tv vector-columns x
// (256 / tv) rows and walks its row slab; per-channel constants are loaded once per thread.
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&f)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kVecU = 4;  // rows in flight per thread

// ring test on the row index within an image (rr < Hp*Wp < 65536): floor(rr / Wp) by reciprocal multiplication
struct RingTest {
  int Hp, Wp, R; unsigned inv;
  __device__ __forceinline__ RingTest(int hp, int wp) : Hp(hp), Wp(wp), R(hp * wp), inv(wp > 0 ? (unsigned)((0x100000000ull + wp - 1) / wp) : 0u) {}
  __device__ __forceinline__ int start(long long m) const { return R > 0 ? (int)(m % R) : 0; }
  __device__ __forceinline__ int advance(int rr, int step) const { rr += step; return R > 0 ? (rr >= R ? rr % R : rr) : 0; }
  __device__ __forceinline__ bool ring(int rr) const {
    if (R <= 0) return false;
    const int y = (int)__umulhi((unsigned)rr, inv), x = rr - y * Wp;
    return y == 0 || y == Hp - 1 || x == 0 || x == Wp - 1;
  }
};

// a raw 8-element packet as loaded from memory (converted to fp32 only when consumed: half the registers in flight)
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 a; };
__device__ __forceinline__ void ldraw(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) { r.a = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&f)[8]) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&f)[8]) {
  const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

template <int MODE, typename TX, typename TD>
__global__ void __launch_bounds__(256, 2) colsum_vec_kernel(const ColDev p, int tv) {
  // slabs are short (<= 64 rows per thread): fp32 partials per thread, doubles only across threads / CTAs
  constexpr int NS = MODE == 0 ? 2 : (MODE == 1 ? 3 : 1);
  constexpr int U = (MODE == 1 && sizeof(TX) + sizeof(TD) > 4) ? 2 : kVecU;
  __shared__ double red[256][8];
  const TX* X = static_cast<const TX*>(p.X);
  const TD* D = static_cast<const TD*>(p.D);
  const int rpi = 256 / tv;
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = (blockIdx.x * tv + vx) * 8;
  const bool active = c < p.C && ry < rpi;
  const long long r_begin = (long long)blockIdx.y * p.rows_per_slab;
  const long long r_end = min(p.m_total, r_begin + p.rows_per_slab);
  float part[NS][8];
#pragma unroll
  for (int j = 0; j < NS; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) part[j][i] = 0.f;
  if (active) {
    float sc[8], sh[8], al[8], mean[8], rstd[8];
    if (MODE == 1) {
      const int fs = p.fold_stride;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc[i] = p.fold[c + i]; sh[i] = p.fold[fs + c + i]; al[i] = p.fold[2 * fs + c + i];
        mean[i] = p.fold[3 * fs + c + i]; rstd[i] = p.fold[4 * fs + c + i];
      }
    }
    const RingTest rt(p.Hp, p.Wp);
    int rr0 = rt.start(r_begin + ry);
    // BN2 backward reductions of the bf16 walk on packed fp32 pairs (8 instead of ~14 instructions per element: the
    // generic loop below ran issue-bound at 54 % of the HBM bandwidth on dense block 1)
    constexpr bool kPacked = MODE == 1 && sizeof(TX) == 2 && sizeof(TD) == 2;
    float2 q0[4], q1[4], q2[4], sc2[4], sh2[4], al2[4], rs2[4], nm2[4];
    if (kPacked) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q0[i] = q1[i] = q2[i] = make_float2(0.f, 0.f);
        sc2[i] = make_float2(sc[2 * i], sc[2 * i + 1]); sh2[i] = make_float2(sh[2 * i], sh[2 * i + 1]);
        al2[i] = make_float2(al[2 * i], al[2 * i + 1]); rs2[i] = make_float2(rstd[2 * i], rstd[2 * i + 1]);
        nm2[i] = make_float2(-mean[2 * i] * rstd[2 * i], -mean[2 * i + 1] * rstd[2 * i + 1]);
      }
    }
    for (long long m0 = r_begin + ry; m0 < r_end; m0 += (long long)rpi * U) {
      Raw8<TX> xr[U];
      Raw8<TD> dr[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long m = m0 + (long long)u * rpi;
        ok[u] = m < r_end && !rt.ring(rr0);
        rr0 = rt.advance(rr0, rpi);
        if (ok[u]) {
          ldraw(X + m * (long long)p.ldx + p.xcol0 + c, xr[u]);
          if (MODE == 1) ldraw(D + m * (long long)p.ldd + p.dcol0 + c, dr[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        if (kPacked) {
          uint32_t xw[4], dw[4];
          memcpy(xw, &xr[u], 16);
          memcpy(dw, &dr[u], 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 x = make_float2(__uint_as_float(xw[i] << 16), __uint_as_float(xw[i] & 0xffff0000u));
            const float2 d = make_float2(__uint_as_float(dw[i] << 16), __uint_as_float(dw[i] & 0xffff0000u));
            const float2 y = __ffma2_rn(x, sc2[i], sh2[i]);
            const float2 g = __fmul2_rn(d, make_float2(y.x >= 0.f ? 1.f : al2[i].x, y.y >= 0.f ? 1.f : al2[i].y));
            q0[i] = __fadd2_rn(q0[i], g);
            q1[i] = __ffma2_rn(g, __ffma2_rn(x, rs2[i], nm2[i]), q1[i]);
            q2[i] = __ffma2_rn(d, make_float2(fminf(y.x, 0.f), fminf(y.y, 0.f)), q2[i]);
          }
          continue;
        }
        float x[8], d[8];
        unpack8(xr[u], x);
        if (MODE == 1) unpack8(dr[u], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (MODE == 0) {
            part[0][i] += x[i];
            part[1][i] = fmaf(x[i], x[i], part[1][i]);
          } else if (MODE == 1) {
            const float y = fmaf(x[i], sc[i], sh[i]);
            const float g = y >= 0.f ? d[i] : d[i] * al[i];
            part[0][i] += g;
            part[1][i] = fmaf(g, (x[i] - mean[i]) * rstd[i], part[1][i]);
            part[2][i] = fmaf(d[i], fminf(y, 0.f), part[2][i]);
          } else {
            part[0][i] += x[i];
          }
        }
      }
    }
    if (kPacked) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        part[0][2 * i] = q0[i].x; part[0][2 * i + 1] = q0[i].y;
        part[1][2 * i] = q1[i].x; part[1][2 * i + 1] = q1[i].y;
        if (NS > 2) { part[NS - 1][2 * i] = q2[i].x; part[NS - 1][2 * i + 1] = q2[i].y; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = (double)part[j][i];
    __syncthreads();
    if (active && ry == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double s = 0.0;
        for (int r = 0; r < rpi; ++r) s += red[r * tv + vx][i];
        p.dst[(size_t)blockIdx.y * p.slot_stride + (size_t)j * p.sum_stride + c + i] = s;
      }
    }
  }
}

template <typename TX, typename TD, typename TO>
__global__ void __launch_bounds__(256, 2) bnact_bwd_apply_vec_kernel(const BnBwdDev p, int tv, int rows_per_slab) {
  constexpr int U = sizeof(TX) + sizeof(TD) + sizeof(TO) > 6 ? 2 : kVecU;
  const int rpi = 256 / tv;
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = (blockIdx.x * tv + vx) * 8;
  if (c >= p.C || ry >= rpi) return;
  const TX* X = static_cast<const TX*>(p.X);
  const TD* D = static_cast<const TD*>(p.D);
  TO* O = static_cast<TO*>(p.dX);
  const long long r_begin = (long long)blockIdx.y * rows_per_slab;
  const long long r_end = min(p.m_total, r_begin + rows_per_slab);
  const int fs = p.fold_stride;
  float sc[8], sh[8], al[8], mean[8], rstd[8], mg[8], mgx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = p.fold[c + i]; sh[i] = p.fold[fs + c + i]; al[i] = p.fold[2 * fs + c + i];
    mean[i] = p.fold[3 * fs + c + i]; rstd[i] = p.fold[4 * fs + c + i];
    mg[i] = (float)(p.sums[c + i] / p.count); mgx[i] = (float)(p.sums[p.C + c + i] / p.count);
  }
  const RingTest rt(p.Hp, p.Wp);
  int rr0 = rt.start(r_begin + ry);
  for (long long m0 = r_begin + ry; m0 < r_end; m0 += (long long)rpi * U) {
    Raw8<TX> xr[U];
    Raw8<TD> dr[U];
    Raw8<TO> orr[U];
    int kind[U];  // 0 skip, 1 ring, 2 interior
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + (long long)u * rpi;
      kind[u] = m >= r_end ? 0 : (rt.ring(rr0) ? 1 : 2);
      rr0 = rt.advance(rr0, rpi);
      if (kind[u] == 2) {
        ldraw(X + m * (long long)p.ldx + p.xcol0 + c, xr[u]);
        ldraw(D + m * (long long)p.ldd + p.dcol0 + c, dr[u]);
        if (p.accumulate) ldraw(O + m * (long long)p.lddx + p.dxcol0 + c, orr[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + (long long)u * rpi;
      if (kind[u] == 0) continue;
      float v[8];
      if (kind[u] == 1) {
        if (p.accumulate) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      } else {
        float x[8], d[8], o[8];
        unpack8(xr[u], x);
        unpack8(dr[u], d);
        if (p.accumulate) unpack8(orr[u], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = fmaf(x[i], sc[i], sh[i]);
          const float g = y >= 0.f ? d[i] : d[i] * al[i];
          v[i] = sc[i] * (g - mg[i] - (x[i] - mean[i]) * rstd[i] * mgx[i]);
          if (p.accumulate) v[i] += o[i];
        }
      }
      st8<TO>(O + m * (long long)p.lddx + p.dxcol0 + c, v);
    }
  }
}

// The bf16 form of the pass above, written for the instruction budget: ncu on the generic kernel showed 28 instructions
// per element at IPC 2.0 and 25 % occupancy - issue-bound at 56 % of the HBM bandwidth, not memory-bound.  Here
//     dx = d * (y >= 0 ? sc : sc * alpha) + (x * B + A),   B = -sc * rstd * mean(g xhat),  A = -sc * mean(g) - B * mean
// is evaluated on packed fp32 pairs (one FFMA2 each for y, x * B + A and the final product: 6 instructions per element),
// with five constants per channel instead of seven.  (x at bf16 precision: folding `mean` into A costs nothing.)
__global__ void __launch_bounds__(256, 2) bnact_bwd_apply16_kernel(const BnBwdDev p, int tv, int rows_per_slab) {
  typedef __nv_bfloat16 bf;
  constexpr int U = kVecU;
  const int rpi = 256 / tv;
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = (blockIdx.x * tv + vx) * 8;
  if (c >= p.C || ry >= rpi) return;
  const bf* X = static_cast<const bf*>(p.X);
  const bf* D = static_cast<const bf*>(p.D);
  bf* O = static_cast<bf*>(p.dX);
  const long long r_begin = (long long)blockIdx.y * rows_per_slab;
  const long long r_end = min(p.m_total, r_begin + rows_per_slab);
  const int fs = p.fold_stride;
  float2 sc[4], sh[4], sn[4], B[4], A[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float k[2][5];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ch = c + 2 * i + h;
      const float s = p.fold[ch], al = p.fold[2 * fs + ch], mean = p.fold[3 * fs + ch], rstd = p.fold[4 * fs + ch];
      const float mg = (float)(p.sums[ch] / p.count), mgx = (float)(p.sums[p.C + ch] / p.count);
      const float b = -s * rstd * mgx;
      k[h][0] = s; k[h][1] = p.fold[fs + ch]; k[h][2] = s * al; k[h][3] = b; k[h][4] = fmaf(-b, mean, -s * mg);
    }
    sc[i] = make_float2(k[0][0], k[1][0]); sh[i] = make_float2(k[0][1], k[1][1]); sn[i] = make_float2(k[0][2], k[1][2]);
    B[i] = make_float2(k[0][3], k[1][3]); A[i] = make_float2(k[0][4], k[1][4]);
  }
  const RingTest rt(p.Hp, p.Wp);
  int rr0 = rt.start(r_begin + ry);
  const bf* xp = X + (r_begin + ry) * (long long)p.ldx + p.xcol0 + c;
  const bf* dp = D + (r_begin + ry) * (long long)p.ldd + p.dcol0 + c;
  bf* op = O + (r_begin + ry) * (long long)p.lddx + p.dxcol0 + c;
  const long long xs = (long long)rpi * p.ldx, ds = (long long)rpi * p.ldd, os = (long long)rpi * p.lddx;
  for (long long m0 = r_begin + ry; m0 < r_end; m0 += (long long)rpi * U) {
    uint4 xr[U], dr[U];
    int kind[U];  // 0 skip, 1 ring, 2 interior
#pragma unroll
    for (int u = 0; u < U; ++u) {
      kind[u] = m0 + (long long)u * rpi >= r_end ? 0 : (rt.ring(rr0) ? 1 : 2);
      rr0 = rt.advance(rr0, rpi);
      if (kind[u] == 2) {
        xr[u] = *reinterpret_cast<const uint4*>(xp + u * xs);
        dr[u] = *reinterpret_cast<const uint4*>(dp + u * ds);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kind[u] == 0) continue;
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (kind[u] == 2) {
        const uint32_t xw[4] = {xr[u].x, xr[u].y, xr[u].z, xr[u].w}, dw[4] = {dr[u].x, dr[u].y, dr[u].z, dr[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = make_float2(__uint_as_float(xw[i] << 16), __uint_as_float(xw[i] & 0xffff0000u));
          const float2 d = make_float2(__uint_as_float(dw[i] << 16), __uint_as_float(dw[i] & 0xffff0000u));
          const float2 y = __ffma2_rn(x, sc[i], sh[i]);
          const float2 s = make_float2(y.x >= 0.f ? sc[i].x : sn[i].x, y.y >= 0.f ? sc[i].y : sn[i].y);
          const float2 v = __ffma2_rn(d, s, __ffma2_rn(x, B[i], A[i]));
          const __nv_bfloat162 h = __float22bfloat162_rn(v);
          ow[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        out = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      *reinterpret_cast<uint4*>(op + u * os) = out;
    }
    xp += U * xs; dp += U * ds; op += U * os;
  }
}

// ------------------------------------------------------------------------------------------------
// BN1 + PReLU1 backward of a dense layer (bf16 walk), reduction half.  All BN1s of a block normalise the same channels
// with the same statistics, so   d x = sum_i sc_i g_i  -  [sum_i sc_i mean(g_i)]  -  xhat [sum_i sc_i mean(g_i xhat)]:
// this pass only reduces (sum g, sum g xhat, sum dA min(y,0)) of layer i into per-slab partial sums; the products
// sc_i g_i are never stored - the consumer of a channel's gradient (grad_pull_kernel, train_cnn.cu) re-derives them from
// the layers' input gradients dA_i (bf16, kept per layer) and applies the two accumulated correction scalars.  Round 1
// added sc_i g_i into an fp32 block gradient here: 8 of the 12 bytes per element of the most expensive backward pass.
// ------------------------------------------------------------------------------------------------
struct Bn1FusedDev {
  const __nv_bfloat16* X; int ldx;
  const __nv_bfloat16* D; int ldd;
  const float* fold; int fold_stride; int C;
  long long m_total; int Hp, Wp, rows_per_slab;
  double* parts;   // [slabs][3][C]
};

__global__ void __launch_bounds__(256, 2) bn1_bwd_reduce_vec_kernel(const Bn1FusedDev p, int tv) {
  constexpr int U = 4;
  __shared__ double red[256][8];
  const int rpi = 256 / tv;
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = (blockIdx.x * tv + vx) * 8;
  const bool active = c < p.C && ry < rpi;
  const long long r_begin = (long long)blockIdx.y * p.rows_per_slab;
  const long long r_end = min(p.m_total, r_begin + p.rows_per_slab);
  float part[3][8];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) part[j][i] = 0.f;
  if (active) {
    const int fs = p.fold_stride;
    float sc[8], sh[8], al[8], mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = p.fold[c + i]; sh[i] = p.fold[fs + c + i]; al[i] = p.fold[2 * fs + c + i];
      mean[i] = p.fold[3 * fs + c + i]; rstd[i] = p.fold[4 * fs + c + i];
    }
    const RingTest rt(p.Hp, p.Wp);
    int rr0 = rt.start(r_begin + ry);
    // packed fp32 pairs, as in colsum_vec_kernel<1>
    float2 q0[4], q1[4], q2[4], sc2[4], sh2[4], al2[4], rs2[4], nm2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q0[i] = q1[i] = q2[i] = make_float2(0.f, 0.f);
      sc2[i] = make_float2(sc[2 * i], sc[2 * i + 1]); sh2[i] = make_float2(sh[2 * i], sh[2 * i + 1]);
      al2[i] = make_float2(al[2 * i], al[2 * i + 1]); rs2[i] = make_float2(rstd[2 * i], rstd[2 * i + 1]);
      nm2[i] = make_float2(-mean[2 * i] * rstd[2 * i], -mean[2 * i + 1] * rstd[2 * i + 1]);
    }
    for (long long m0 = r_begin + ry; m0 < r_end; m0 += (long long)rpi * U) {
      Raw8<__nv_bfloat16> xr[U], dr[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long m = m0 + (long long)u * rpi;
        ok[u] = m < r_end && !rt.ring(rr0);
        rr0 = rt.advance(rr0, rpi);
        if (ok[u]) {
          ldraw(p.X + m * (long long)p.ldx + c, xr[u]);
          ldraw(p.D + m * (long long)p.ldd + c, dr[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const uint32_t xw[4] = {xr[u].a.x, xr[u].a.y, xr[u].a.z, xr[u].a.w}, dw[4] = {dr[u].a.x, dr[u].a.y, dr[u].a.z, dr[u].a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = make_float2(__uint_as_float(xw[i] << 16), __uint_as_float(xw[i] & 0xffff0000u));
          const float2 d = make_float2(__uint_as_float(dw[i] << 16), __uint_as_float(dw[i] & 0xffff0000u));
          const float2 y = __ffma2_rn(x, sc2[i], sh2[i]);
          const float2 g = __fmul2_rn(d, make_float2(y.x >= 0.f ? 1.f : al2[i].x, y.y >= 0.f ? 1.f : al2[i].y));
          q0[i] = __fadd2_rn(q0[i], g);
          q1[i] = __ffma2_rn(g, __ffma2_rn(x, rs2[i], nm2[i]), q1[i]);
          q2[i] = __ffma2_rn(d, make_float2(fminf(y.x, 0.f), fminf(y.y, 0.f)), q2[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      part[0][2 * i] = q0[i].x; part[0][2 * i + 1] = q0[i].y;
      part[1][2 * i] = q1[i].x; part[1][2 * i + 1] = q1[i].y;
      part[2][2 * i] = q2[i].x; part[2][2 * i + 1] = q2[i].y;
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = (double)part[j][i];
    __syncthreads();
    if (active && ry == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double sum = 0.0;
        for (int r = 0; r < rpi; ++r) sum += red[r * tv + vx][i];
        p.parts[((size_t)blockIdx.y * 3 + j) * p.C + c + i] = sum;
      }
    }
  }
}

// out[j * out_stride + c] = sum over the slots of parts[slot][j][c], slots added in a fixed order: a block owns 16 columns,
// its 16 slices each add every 16th slot (128-byte coalesced reads), then the slice sums are added in order
__global__ void __launch_bounds__(256) parts_reduce_kernel(const double* __restrict__ parts, int n_slots, int ns, int C,
                                                           double* __restrict__ out, int out_stride) {
  __shared__ double red[16][17];
  const int col = threadIdx.x & 15, slice = threadIdx.x >> 4;
  const int q = blockIdx.x * 16 + col;   // flattened (sum j, column c)
  const int total = ns * C;
  double acc = 0.0;
  if (q < total) {
    int s = slice;
    for (; s + 16 * 7 < n_slots; s += 16 * 8) {   // 8 independent loads in flight, added in slot order
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = parts[(size_t)(s + 16 * u) * total + q];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; s < n_slots; s += 16) acc += parts[(size_t)s * total + q];
  }
  red[slice][col] = acc;
  __syncthreads();
  if (slice == 0 && q < total) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += red[k][col];
    out[(size_t)(q / C) * out_stride + q % C] = t;
  }
}

template <typename TX, typename TO>
__global__ void __launch_bounds__(256) bnact_fwd_vec_kernel(const TX* __restrict__ X, int ldx, int xcol0,
                                                            const float* __restrict__ fold, int fold_stride, int C,
                                                            long long m_total, int Hp, int Wp, TO* __restrict__ out, int ldo,
                                                            int ocol0, int tv, int rows_per_slab) {
  // a block is tv vector-columns x (256 / tv) rows walking a row slab: no per-element index division, the ring test is
  // incremental, the per-channel constants are loaded once per thread
  const int rpi = 256 / tv;
  const int vx = threadIdx.x % tv, ry = threadIdx.x / tv;
  const int c = (blockIdx.x * tv + vx) * 8;
  if (c >= C || ry >= rpi) return;
  const long long r_begin = (long long)blockIdx.y * rows_per_slab;
  const long long r_end = min(m_total, r_begin + rows_per_slab);
  float sc[8], sh[8], al[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = fold[c + i]; sh[i] = fold[fold_stride + c + i]; al[i] = fold[2 * fold_stride + c + i]; }
  const RingTest rt(Hp, Wp);
  int rr0 = rt.start(r_begin + ry);
  for (long long m0 = r_begin + ry; m0 < r_end; m0 += (long long)rpi * kVecU) {
    Raw8<TX> xr[kVecU];
    int kind[kVecU];
#pragma unroll
    for (int u = 0; u < kVecU; ++u) {
      const long long m = m0 + (long long)u * rpi;
      kind[u] = m >= r_end ? 0 : (rt.ring(rr0) ? 1 : 2);
      rr0 = rt.advance(rr0, rpi);
      if (kind[u] == 2) ldraw(X + m * (long long)ldx + xcol0 + c, xr[u]);
    }
#pragma unroll
    for (int u = 0; u < kVecU; ++u) {
      if (kind[u] == 0) continue;
      const long long m = m0 + (long long)u * rpi;
      float v[8];
      if (kind[u] == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      } else {
        unpack8(xr[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = prelu(fmaf(v[i], sc[i], sh[i]), al[i]);
      }
      st8<TO>(out + m * (long long)ldo + ocol0 + c, v);
    }
  }
}

// batch statistics -> fold [scale | shift | alpha | mean | rstd], running-stat update (momentum, unbiased var)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ alpha, float eps,
                                   float momentum, float* running_mean, float* running_var, float* __restrict__ fold) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  fold[c] = sc;
  fold[C + c] = beta[c] - (float)mean * sc;
  fold[2 * C + c] = alpha ? alpha[c] : 1.f;
  fold[3 * C + c] = (float)mean;
  fold[4 * C + c] = rstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// BN + PReLU backward, elementwise half:  dX (+)= scale * (g - sum_g/M - xhat * sum_gxhat/M)
template <typename TX, typename TD, typename TO>
__global__ void bnact_bwd_apply_kernel(const BnBwdDev p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.m_total * p.C) return;
  const int c = (int)(idx % p.C);
  const long long m = idx / p.C;
  TO* dst = static_cast<TO*>(p.dX) + m * (long long)p.lddx + p.dxcol0 + c;
  if (is_ring(m, p.Hp, p.Wp)) {
    if (!p.accumulate) *dst = from_f32<TO>(0.f);
    return;
  }
  const int fs = p.fold_stride;
  const float sc = p.fold[c], sh = p.fold[fs + c], al = p.fold[2 * fs + c], mean = p.fold[3 * fs + c], rstd = p.fold[4 * fs + c];
  const float x = to_f32<TX>(static_cast<const TX*>(p.X)[m * (long long)p.ldx + p.xcol0 + c]);
  const float d = to_f32<TD>(static_cast<const TD*>(p.D)[m * (long long)p.ldd + p.dcol0 + c]);
  const float y = fmaf(x, sc, sh);
  const float g = y >= 0.f ? d : d * al;
  const float mg = (float)(p.sums[c] / p.count), mgx = (float)(p.sums[p.C + c] / p.count);
  const float v = sc * (g - mg - (x - mean) * rstd * mgx);
  *dst = from_f32<TO>(p.accumulate ? to_f32<TO>(*dst) + v : v);
}

// parameter gradients of a BN + PReLU pair from the reductions: dgamma = sum g*xhat, dbeta = sum g, dalpha
__global__ void bnact_param_grads_kernel(const double* __restrict__ sums, int C, float* dgamma, float* dbeta, float* dalpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dgamma) dgamma[c] += (float)sums[C + c];
  if (dbeta) dbeta[c] += (float)sums[c];
  if (dalpha) dalpha[c] += (float)sums[2 * C + c];
}

__global__ void add_cols_kernel(const double* __restrict__ sums, int C, float* dst) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dst[c] += (float)sums[c];
}

// out[m, c] = PReLU(BN(x[m, c])) (ring rows -> 0): materialised activation for the few places that need it
template <typename TX, typename TO>
__global__ void bnact_fwd_kernel(const TX* __restrict__ X, int ldx, int xcol0, const float* __restrict__ fold, int fold_stride,
                                 int C, long long m_total, int Hp, int Wp, TO* __restrict__ out, int ldo, int ocol0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m_total * C) return;
  const int c = (int)(idx % C);
  const long long m = idx / C;
  float v = 0.f;
  if (!is_ring(m, Hp, Wp))
    v = prelu(fmaf(to_f32<TX>(X[m * (long long)ldx + xcol0 + c]), fold[c], fold[fold_stride + c]), fold[2 * fold_stride + c]);
  out[m * (long long)ldo + ocol0 + c] = from_f32<TO>(v);
}

// ------------------------------------------------------------------------------------------------
// pooling forward / backward on ringed maps
// ------------------------------------------------------------------------------------------------
// AvgPool2d(3,2) of the un-ringed stem map act(z0) [n,Hs,Ws,C] -> ringed [n,H+2,W+2,ld] channels [0,C)
template <typename TO>
__global__ void stem_pool_fwd_kernel(const float* __restrict__ z, const float* __restrict__ fold, int Hs, int Ws, int C,
                                     TO* __restrict__ blk, int ld, int H, int W, long long total) {
  // one thread = one pooled pixel x 4 channels (float4 loads of the 3x3 window); total counts those quads
  const long long idx64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx64 >= total) return;
  const unsigned idx = (unsigned)idx64;   // the launcher bounds total below 2^31: 32-bit divisions
  const unsigned cq = (unsigned)C >> 2;
  const int c = (int)(idx % cq) * 4;
  unsigned r = idx / cq;
  const int x = (int)(r % (unsigned)W); r /= (unsigned)W;
  const int y = (int)(r % (unsigned)H);
  const int n = (int)(r / (unsigned)H);
  const float4 sc = *reinterpret_cast<const float4*>(fold + c), sh = *reinterpret_cast<const float4*>(fold + C + c),
               al = *reinterpret_cast<const float4*>(fold + 2 * C + c);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const float4 v = *reinterpret_cast<const float4*>(z + (((size_t)n * Hs + 2 * y + dy) * Ws + 2 * x + dx) * C + c);
      s0 += prelu(fmaf(v.x, sc.x, sh.x), al.x);
      s1 += prelu(fmaf(v.y, sc.y, sh.y), al.y);
      s2 += prelu(fmaf(v.z, sc.z, sh.z), al.z);
      s3 += prelu(fmaf(v.w, sc.w, sh.w), al.w);
    }
  TO* dst = blk + ((size_t)n * (H + 2) * (W + 2) + (size_t)(y + 1) * (W + 2) + x + 1) * ld + c;
  dst[0] = from_f32<TO>(s0 / 9.0f); dst[1] = from_f32<TO>(s1 / 9.0f); dst[2] = from_f32<TO>(s2 / 9.0f); dst[3] = from_f32<TO>(s3 / 9.0f);
}

// gradient of the above w.r.t. the activated stem map: dA[n,oy,ox,c] = (1/9) sum of dP over the windows holding (oy,ox)
template <typename TO>
__global__ void stem_pool_bwd_kernel(const float* __restrict__ dblk, int ld, int H, int W, int C, TO* __restrict__ dA,
                                     int Hs, int Ws, long long total) {
  // one thread = one stem pixel x 4 channels; window py covers rows 2py .. 2py+2
  const long long idx64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx64 >= total) return;
  const unsigned idx = (unsigned)idx64;   // the launcher bounds total below 2^31: 32-bit divisions
  const unsigned cq = (unsigned)C >> 2;
  const int c = (int)(idx % cq) * 4;
  unsigned r = idx / cq;
  const int ox = (int)(r % (unsigned)Ws); r /= (unsigned)Ws;
  const int oy = (int)(r % (unsigned)Hs);
  const int n = (int)(r / (unsigned)Hs);
  const int py_lo = oy >= 2 ? (oy - 1) >> 1 : 0, py_hi = min(H - 1, oy >> 1);
  const int px_lo = ox >= 2 ? (ox - 1) >> 1 : 0, px_hi = min(W - 1, ox >> 1);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int py = py_lo; py <= py_hi; ++py)
    for (int px = px_lo; px <= px_hi; ++px) {
      const float4 v = *reinterpret_cast<const float4*>(dblk + ((size_t)n * (H + 2) * (W + 2) + (size_t)(py + 1) * (W + 2) + px + 1) * ld + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  TO* dst = dA + (((size_t)n * Hs + oy) * Ws + ox) * C + c;
  if (sizeof(TO) == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(s.x / 9.0f, s.y / 9.0f, s.z / 9.0f, s.w / 9.0f);
  } else {   // bf16 stem gradient map (bf16 walk): half the bytes of the three dense passes over it
    const __nv_bfloat162 lo = __floats2bfloat162_rn(s.x / 9.0f, s.y / 9.0f), hi = __floats2bfloat162_rn(s.z / 9.0f, s.w / 9.0f);
    *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}

// the same pair on all-bf16 maps (bf16 walk), one thread = one pixel x 8 channels (16-byte loads / stores)
__global__ void __launch_bounds__(256) stem_pool16_fwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ fold, int Hs,
                                                             int Ws, int C, __nv_bfloat16* __restrict__ blk, int ld, int H, int W,
                                                             long long total) {
  // grid (column octets of a pooled row, pooled row, image): one small division per thread instead of three
  (void)total;
  const unsigned cv = (unsigned)C >> 3;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)W * cv) return;
  const int x = (int)(t / cv);
  const int c = (int)(t - (unsigned)x * cv) * 8;
  const int y = blockIdx.y, n = blockIdx.z;
  float sc[8], sh[8], al[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = fold[c + i]; sh[i] = fold[C + c + i]; al[i] = fold[2 * C + c + i]; s[i] = 0.f; }
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      float v[8];
      ld8<__nv_bfloat16>(z + (((size_t)n * Hs + 2 * y + dy) * Ws + 2 * x + dx) * C + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += prelu(fmaf(v[i], sc[i], sh[i]), al[i]);
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] *= (1.0f / 9.0f);   // bf16 path: a product (<= 1 fp32 ulp from the quotient, 10x fewer instructions)
  st8<__nv_bfloat16>(blk + ((size_t)n * (H + 2) * (W + 2) + (size_t)(y + 1) * (W + 2) + x + 1) * ld + c, s);
}

__global__ void __launch_bounds__(256) stem_pool16_bwd_kernel(const __nv_bfloat16* __restrict__ dblk, int ld, int H, int W, int C,
                                                             __nv_bfloat16* __restrict__ dA, int Hs, int Ws, long long total) {
  // grid (column octets of a stem row, stem row, image)
  (void)total;
  const unsigned cv = (unsigned)C >> 3;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)Ws * cv) return;
  const int ox = (int)(t / cv);
  const int c = (int)(t - (unsigned)ox * cv) * 8;
  const int oy = blockIdx.y, n = blockIdx.z;
  const int py_lo = oy >= 2 ? (oy - 1) >> 1 : 0, py_hi = min(H - 1, oy >> 1);
  const int px_lo = ox >= 2 ? (ox - 1) >> 1 : 0, px_hi = min(W - 1, ox >> 1);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int py = py_lo; py <= py_hi; ++py)
    for (int px = px_lo; px <= px_hi; ++px) {
      float v[8];
      ld8<__nv_bfloat16>(dblk + ((size_t)n * (H + 2) * (W + 2) + (size_t)(py + 1) * (W + 2) + px + 1) * ld + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += v[i];
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] *= (1.0f / 9.0f);   // bf16 path: a product (<= 1 fp32 ulp from the quotient, 10x fewer instructions)
  st8<__nv_bfloat16>(dA + (((size_t)n * Hs + oy) * Ws + ox) * C + c, s);
}

// AvgPool2d(2,2) backward: dA[ringed H x W rows, C] = 0.25 * dP[ringed H2 x W2 parent] (0 where the floor cropped)
template <typename T>
__global__ void pool2_bwd_kernel(const T* __restrict__ dP, int H2, int W2, int C, T* __restrict__ dA, int H, int W,
                                 long long total) {
  // one thread = one ringed pixel x 8 channels; total counts those octets.  Index arithmetic in 32 bits (the launcher keeps
  // rows * C / 8 below 2^31): with 64-bit divisions this pass was issue-bound at 2.0-2.8 TB/s.
  const long long idx64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx64 >= total) return;
  const unsigned idx = (unsigned)idx64;
  const unsigned cv = (unsigned)C >> 3;
  const unsigned Wp = W + 2, Hp = H + 2;
  const int c = (int)(idx % cv) * 8;
  unsigned r = idx / cv;
  const int xx = (int)(r % Wp); r /= Wp;
  const int yy = (int)(r % Hp);
  const int n = (int)(r / Hp);
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  const int y = yy - 1, x = xx - 1;
  if (y >= 0 && y < 2 * H2 && x >= 0 && x < 2 * W2) {
    ld8<T>(dP + ((size_t)n * (H2 + 2) * (W2 + 2) + (size_t)(y / 2 + 1) * (W2 + 2) + x / 2 + 1) * C + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= 0.25f;
  }
  st8<T>(dA + (size_t)idx64 * 8, v);
}

// global average pool backward: dA[ringed rows, C] = dGap[n, c] / (H*W) on interior rows
template <typename TO>
__global__ void gap_bwd_kernel(const float* __restrict__ dGap, int C, TO* __restrict__ dA, int H, int W, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long m = idx / C;
  const int R = (H + 2) * (W + 2);
  const int n = (int)(m / R);
  dA[idx] = from_f32<TO>(is_ring(m, H + 2, W + 2) ? 0.f : dGap[(size_t)n * C + c] / (float)(H * W));
}

// ------------------------------------------------------------------------------------------------
// dropout: x *= mask / (1 - p), mask from a counter-based generator keyed by (seed, stream, element)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint64_t z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

template <typename T>
__global__ void dropout_kernel(T* X, int ld, int col0, int C, long long m_total, unsigned long long seed0,
                               unsigned long long stream_id, float p, const unsigned long long* seed_off) {
  const unsigned long long seed = seed_with_offset(seed0, seed_off);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m_total * C) return;
  const int c = (int)(idx % C);
  const long long m = idx / C;
  const uint32_t r = mix32(seed * 0x100000001b3ull + stream_id * 0x9e3779b97f4a7c15ull + (unsigned long long)idx);
  const bool keep = (r >> 8) * (1.0f / 16777216.0f) >= p;
  T* v = X + m * (long long)ld + col0 + c;
  *v = from_f32<T>(keep ? to_f32<T>(*v) / (1.f - p) : 0.f);
}

// ------------------------------------------------------------------------------------------------
// stem convolution for training: raw conv0 output (no bias fold) and its weight gradient, both hit-driven
// ------------------------------------------------------------------------------------------------
// z[n,oy,ox,c] = b[c] + sum over hits: one thread per (hit, output position) pair walks the 64 channels
__global__ void stem_fill_bias_kernel(float* z, const float* __restrict__ bias, int C, long long total4) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 per thread (C % 4 == 0)
  if (idx < total4) {   // the bias lives in the parameter arena: no alignment guarantee, scalar loads
    const int c = (int)(((unsigned)idx * 4u) % (unsigned)C);   // the launcher bounds idx * 4 below 2^32: 32-bit modulo
    reinterpret_cast<float4*>(z)[idx] = make_float4(__ldg(bias + c), __ldg(bias + c + 1), __ldg(bias + c + 2), __ldg(bias + c + 3));
  }
}

// pixels NCHW fp32; for every non-zero input value scatter v*w into the <=16 outputs it reaches (atomic: training path).
// Persistent grid, one warp per 32 consecutive values per step, lanes then share the channel loop.  WGRAD: the filter
// gradient is accumulated in shared memory (cin*49*C <= 9408 floats) and flushed once per CTA - global atomics on the
// 9.4 k filter entries from every hit of the batch serialise otherwise.
constexpr int kStemW = 3 * 49 * 64;
template <bool WGRAD, typename TDZ = float>
__global__ void __launch_bounds__(256) stem_conv_scatter_kernel(const float* __restrict__ pixels, int cin, int H, int W, int Hs,
                                                                int Ws, const float* __restrict__ w /*[cin*49][C]*/, int C,
                                                                float* __restrict__ z, const TDZ* __restrict__ dz,
                                                                float* __restrict__ dw, unsigned total) {
  __shared__ float sdw[WGRAD ? kStemW : 1];
  const int nw = cin * 49 * C;
  if (WGRAD) {
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sdw[i] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned step = gridDim.x * 8u * 32u;
  for (unsigned base = (blockIdx.x * 8u + warp) * 32u; base < total; base += step) {
    const unsigned idx = base + lane;
    float v = 0.f;
    if (idx < total) v = __ldg(pixels + idx);
    unsigned nz = __ballot_sync(0xffffffffu, v != 0.f);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const float val = __shfl_sync(0xffffffffu, v, src);
      const unsigned e = base + src;
      const int x = (int)(e % (unsigned)W);
      unsigned r = e / (unsigned)W;
      const int y = (int)(r % (unsigned)H); r /= (unsigned)H;
      const int c = (int)(r % (unsigned)cin);
      const int n = (int)(r / (unsigned)cin);
      // outputs (oy, ox) with 2*oy - 3 + ky == y, ky in [0,7)
      for (int ky = (y + 3) & 1; ky < 7; ky += 2) {
        const int oy = (y + 3 - ky) >> 1;
        if (oy < 0 || oy >= Hs) continue;
        for (int kx = (x + 3) & 1; kx < 7; kx += 2) {
          const int ox = (x + 3 - kx) >> 1;
          if (ox < 0 || ox >= Ws) continue;
          const size_t o = (((size_t)n * Hs + oy) * Ws + ox) * C;
          const int wi = ((c * 7 + ky) * 7 + kx) * C;
          for (int ch = lane; ch < C; ch += 32) {
            if (!WGRAD) atomicAdd(z + o + ch, val * __ldg(w + wi + ch));
            else atomicAdd(&sdw[wi + ch], val * to_f32<TDZ>(dz[o + ch]));
          }
        }
      }
    }
    if (base + step < base) break;   // unsigned wrap-around guard
  }
  if (WGRAD) {
    __syncthreads();
    for (int i = threadIdx.x; i < nw; i += blockDim.x) {
      const float a = sdw[i];
      if (a != 0.f) atomicAdd(dw + i, a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------------
static int slabs_for(long long rows, int* rows_per_slab) {
  int slabs = (int)ceil_div_ll(rows, 256);
  if (slabs > 592) slabs = 592;
  if (slabs < 1) slabs = 1;
  *rows_per_slab = (int)ceil_div_ll(rows, slabs);
  return (int)ceil_div_ll(rows, *rows_per_slab);
}

static inline int colsum_ns_of(int mode) { return mode == 0 ? 2 : (mode == 1 ? 3 : 1); }

// per-slab partial column sums: dst[slab * slot_stride + j * sum_stride + c]; *n_slabs = slabs written.
// one_slab forces a single slab (the result is then final).  x_bf16 / d_bf16: element types of X / D.
static int colsums_launch(int mode, const void* X, bool x_bf16, int ldx, int xcol0, const void* D, bool d_bf16, int ldd, int dcol0,
                          const float* fold, int fold_stride, int C, long long m_total, int ring_hp, int ring_wp, double* dst,
                          long long slot_stride, int sum_stride, bool one_slab, int* n_slabs, cudaStream_t stream) {
  ColDev p{};
  p.X = X; p.ldx = ldx; p.xcol0 = xcol0; p.D = D; p.ldd = ldd; p.dcol0 = dcol0; p.fold = fold; p.fold_stride = fold_stride;
  p.C = C; p.m_total = m_total; p.Hp = ring_hp; p.Wp = ring_wp; p.dst = dst; p.slot_stride = slot_stride; p.sum_stride = sum_stride;
  int slabs = slabs_for(m_total, &p.rows_per_slab);
  typedef __nv_bfloat16 bf;
  const bool vec = C % 8 == 0 && ldx % 8 == 0 && xcol0 % 8 == 0 && (mode != 1 || (ldd % 8 == 0 && dcol0 % 8 == 0)) &&
                   reinterpret_cast<uintptr_t>(X) % 32 == 0 && (mode != 1 || reinterpret_cast<uintptr_t>(D) % 32 == 0);
  if (vec) {
    const int cv = C / 8;
    const int tv = cv < 32 ? cv : 32;
    const int rpi = 256 / tv;
    const int gx = ceil_div(cv, tv);
    // short slabs: one slab = 16 row-steps of a CTA, at most ~4 CTAs per SM in total
    long long vslabs = ceil_div_ll(m_total, (long long)rpi * kVecU * 4);
    const long long cap = (long long)(148 * 2 / gx > 1 ? 148 * 2 / gx : 1);
    if (vslabs > cap) vslabs = cap;
    if (one_slab) vslabs = 1;
    p.rows_per_slab = (int)ceil_div_ll(m_total, vslabs);
    vslabs = ceil_div_ll(m_total, p.rows_per_slab);
    *n_slabs = (int)vslabs;
    dim3 vgrid(gx, (unsigned)vslabs);
#define TCVN_COLSUM_V(MODE)                                                                              \
  do {                                                                                                   \
    if (!x_bf16 && !d_bf16) colsum_vec_kernel<MODE, float, float><<<vgrid, 256, 0, stream>>>(p, tv);      \
    else if (x_bf16 && d_bf16) colsum_vec_kernel<MODE, bf, bf><<<vgrid, 256, 0, stream>>>(p, tv);         \
    else if (x_bf16) colsum_vec_kernel<MODE, bf, float><<<vgrid, 256, 0, stream>>>(p, tv);                \
    else colsum_vec_kernel<MODE, float, bf><<<vgrid, 256, 0, stream>>>(p, tv);                            \
  } while (0)
    if (mode == 0) TCVN_COLSUM_V(0);
    else if (mode == 1) TCVN_COLSUM_V(1);
    else TCVN_COLSUM_V(2);
#undef TCVN_COLSUM_V
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }
  {
    const long long cap = (long long)592 * 3 * 256 / ((long long)colsum_ns_of(mode) * C);   // capacity of the parts scratch
    if (slabs > cap) slabs = (int)(cap < 1 ? 1 : cap);
    if (one_slab) slabs = 1;
    p.rows_per_slab = (int)ceil_div_ll(m_total, slabs);
    slabs = (int)ceil_div_ll(m_total, p.rows_per_slab);
  }
  *n_slabs = slabs;
  dim3 grid(ceil_div(C, 32), slabs);
#define TCVN_COLSUM(MODE)                                                                          \
  do {                                                                                             \
    if (!x_bf16 && !d_bf16) colsum_kernel<MODE, float, float><<<grid, 256, 0, stream>>>(p);         \
    else if (x_bf16 && d_bf16) colsum_kernel<MODE, bf, bf><<<grid, 256, 0, stream>>>(p);            \
    else if (x_bf16) colsum_kernel<MODE, bf, float><<<grid, 256, 0, stream>>>(p);                   \
    else colsum_kernel<MODE, float, bf><<<grid, 256, 0, stream>>>(p);                               \
  } while (0)
  if (mode == 0) TCVN_COLSUM(0);
  else if (mode == 1) TCVN_COLSUM(1);
  else TCVN_COLSUM(2);
#undef TCVN_COLSUM
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

static inline int colsum_ns(int mode) { return colsum_ns_of(mode); }

size_t colsum_parts_bytes() { return (size_t)592 * 3 * 256 * sizeof(double) + 4096; }

// per-slab partial sums only: parts[slab][ns][C] (the consumer adds the slabs in a fixed order)
int colsums_parts(int mode, const void* X, bool x_bf16, int ldx, int xcol0, const void* D, bool d_bf16, int ldd, int dcol0,
                  const float* fold, int fold_stride, int C, long long m_total, int ring_hp, int ring_wp, double* parts,
                  int* n_slabs, cudaStream_t stream) {
  *n_slabs = 0;
  if (m_total <= 0 || C <= 0) return TCVN_OK;
  return colsums_launch(mode, X, x_bf16, ldx, xcol0, D, d_bf16, ldd, dcol0, fold, fold_stride, C, m_total, ring_hp, ring_wp, parts,
                        (long long)colsum_ns(mode) * C, C, false, n_slabs, stream);
}

// column sums ASSIGNED to out[j * out_stride + c], bit-reproducible: per-slab partial sums in `parts` (scratch of
// colsum_parts_bytes()) added in a fixed order; parts == nullptr -> one slab writing the result directly (small inputs /
// test hooks).
int colsums_typed(int mode, const void* X, bool x_bf16, int ldx, int xcol0, const void* D, bool d_bf16, int ldd, int dcol0,
                  const float* fold, int fold_stride, int C, long long m_total, int ring_hp, int ring_wp, double* out,
                  int out_stride, double* parts, cudaStream_t stream) {
  if (C <= 0) return TCVN_OK;
  const int ns = colsum_ns(mode);
  if (m_total <= 0) {
    for (int j = 0; j < ns; ++j) TCVN_CUDA(cudaMemsetAsync(out + (size_t)j * out_stride, 0, sizeof(double) * C, stream));
    return TCVN_OK;
  }
  int slabs = 0;
  if (parts == nullptr)
    return colsums_launch(mode, X, x_bf16, ldx, xcol0, D, d_bf16, ldd, dcol0, fold, fold_stride, C, m_total, ring_hp, ring_wp, out,
                          0, out_stride, true, &slabs, stream);
  TCVN_TRY(colsums_launch(mode, X, x_bf16, ldx, xcol0, D, d_bf16, ldd, dcol0, fold, fold_stride, C, m_total, ring_hp, ring_wp,
                          parts, (long long)ns * C, C, false, &slabs, stream));
  parts_reduce_kernel<<<ceil_div(ns * C, 16), 256, 0, stream>>>(parts, slabs, ns, C, out, out_stride);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int colsums_into(int mode, const float* X, int ldx, int xcol0, const float* D, int ldd, int dcol0, const float* fold, int C,
                 long long m_total, int ring_hp, int ring_wp, double* out, int out_stride, double* parts, cudaStream_t stream) {
  return colsums_typed(mode, X, false, ldx, xcol0, D, false, ldd, dcol0, fold, C, C, m_total, ring_hp, ring_wp, out, out_stride,
                       parts, stream);
}

// dX (+)= BN + PReLU backward of D at X (elementwise half); types: 0 = fp32, 1 = bf16
int bnact_bwd_apply_typed(const void* D, bool d_bf16, int ldd, int dcol0, const void* X, bool x_bf16, int ldx, int xcol0,
                          const float* fold, int fold_stride, const double* sums, int C, double count, void* dX, bool o_bf16,
                          int lddx, int dxcol0, bool accumulate, long long m_total, int ring_hp, int ring_wp,
                          cudaStream_t stream) {
  if (m_total <= 0 || C <= 0) return TCVN_OK;
  BnBwdDev p{};
  p.D = D; p.ldd = ldd; p.dcol0 = dcol0; p.X = X; p.ldx = ldx; p.xcol0 = xcol0; p.fold = fold; p.fold_stride = fold_stride;
  p.sums = sums; p.C = C; p.count = count; p.dX = dX; p.lddx = lddx; p.dxcol0 = dxcol0; p.accumulate = accumulate ? 1 : 0;
  p.m_total = m_total; p.Hp = ring_hp; p.Wp = ring_wp;
  typedef __nv_bfloat16 bf;
  const bool vec = C % 8 == 0 && ldx % 8 == 0 && xcol0 % 8 == 0 && ldd % 8 == 0 && dcol0 % 8 == 0 && lddx % 8 == 0 &&
                   dxcol0 % 8 == 0 && reinterpret_cast<uintptr_t>(X) % 32 == 0 && reinterpret_cast<uintptr_t>(D) % 32 == 0 &&
                   reinterpret_cast<uintptr_t>(dX) % 32 == 0;
  if (vec) {
    const int cv = C / 8;
    const int tv = cv < 32 ? cv : 32;
    const int rpi = 256 / tv;
    const int gx = ceil_div(cv, tv);
    long long vslabs = ceil_div_ll(m_total, (long long)rpi * kVecU * 2);
    const long long cap = (long long)(148 * 2 / gx > 1 ? 148 * 2 / gx : 1);
    if (vslabs > cap) vslabs = cap;
    const int rows_per_slab = (int)ceil_div_ll(m_total, vslabs);
    const int slabs = (int)ceil_div_ll(m_total, rows_per_slab);
    dim3 vgrid(gx, slabs);
    if (!x_bf16 && !d_bf16 && !o_bf16) bnact_bwd_apply_vec_kernel<float, float, float><<<vgrid, 256, 0, stream>>>(p, tv, rows_per_slab);
    else if (x_bf16 && d_bf16 && o_bf16 && !accumulate) bnact_bwd_apply16_kernel<<<vgrid, 256, 0, stream>>>(p, tv, rows_per_slab);
    else if (x_bf16 && d_bf16 && o_bf16) bnact_bwd_apply_vec_kernel<bf, bf, bf><<<vgrid, 256, 0, stream>>>(p, tv, rows_per_slab);
    else if (x_bf16 && d_bf16 && !o_bf16) bnact_bwd_apply_vec_kernel<bf, bf, float><<<vgrid, 256, 0, stream>>>(p, tv, rows_per_slab);
    else if (!x_bf16 && d_bf16 && o_bf16) bnact_bwd_apply_vec_kernel<float, bf, bf><<<vgrid, 256, 0, stream>>>(p, tv, rows_per_slab);
    else return fail(TCVN_ERR_UNSUPPORTED, "bnact_bwd_apply: type combination not instantiated");
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }
  const unsigned grid = (unsigned)ceil_div_ll(m_total * C, 256);
  if (!x_bf16 && !d_bf16 && !o_bf16) bnact_bwd_apply_kernel<float, float, float><<<grid, 256, 0, stream>>>(p);
  else if (x_bf16 && d_bf16 && o_bf16) bnact_bwd_apply_kernel<bf, bf, bf><<<grid, 256, 0, stream>>>(p);
  else if (x_bf16 && d_bf16 && !o_bf16) bnact_bwd_apply_kernel<bf, bf, float><<<grid, 256, 0, stream>>>(p);
  else if (!x_bf16 && d_bf16 && o_bf16) bnact_bwd_apply_kernel<float, bf, bf><<<grid, 256, 0, stream>>>(p);
  else return fail(TCVN_ERR_UNSUPPORTED, "bnact_bwd_apply: type combination not instantiated");
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int bnact_fwd_typed(const void* X, bool x_bf16, int ldx, int xcol0, const float* fold, int fold_stride, int C, long long m_total,
                    int ring_hp, int ring_wp, void* out, bool o_bf16, int ldo, int ocol0, cudaStream_t stream) {
  if (m_total <= 0 || C <= 0) return TCVN_OK;
  typedef __nv_bfloat16 bf;
  if (x_bf16 && o_bf16 && C % 8 == 0 && ldx % 8 == 0 && xcol0 % 8 == 0 && ldo % 8 == 0 && ocol0 % 8 == 0 &&
      reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    const int cv = C / 8;
    const int tv = cv < 32 ? cv : 32;
    const int rpi = 256 / tv;
    const int gx = ceil_div(cv, tv);
    long long vslabs = ceil_div_ll(m_total, (long long)rpi * kVecU * 2);
    const long long cap = (long long)(148 * 8 / gx > 1 ? 148 * 8 / gx : 1);
    if (vslabs > cap) vslabs = cap;
    const int rows_per_slab = (int)ceil_div_ll(m_total, vslabs);
    dim3 vgrid(gx, (unsigned)ceil_div_ll(m_total, rows_per_slab));
    bnact_fwd_vec_kernel<bf, bf><<<vgrid, 256, 0, stream>>>(static_cast<const bf*>(X), ldx, xcol0, fold, fold_stride, C, m_total,
                                                            ring_hp, ring_wp, static_cast<bf*>(out), ldo, ocol0, tv, rows_per_slab);
    TCVN_LAUNCH_CHECK();
    return TCVN_OK;
  }
  const unsigned grid = (unsigned)ceil_div_ll(m_total * C, 256);
  if (!x_bf16 && !o_bf16)
    bnact_fwd_kernel<float, float><<<grid, 256, 0, stream>>>(static_cast<const float*>(X), ldx, xcol0, fold, fold_stride, C, m_total,
                                                             ring_hp, ring_wp, static_cast<float*>(out), ldo, ocol0);
  else if (x_bf16 && o_bf16)
    bnact_fwd_kernel<bf, bf><<<grid, 256, 0, stream>>>(static_cast<const bf*>(X), ldx, xcol0, fold, fold_stride, C, m_total, ring_hp,
                                                       ring_wp, static_cast<bf*>(out), ldo, ocol0);
  else return fail(TCVN_ERR_UNSUPPORTED, "bnact_fwd: type combination not instantiated");
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// kinds as tcvn_t_pool; bf16 = element type of the ringed block-side buffers (the stem map z0 and dGap stay fp32)
int pool_typed(int kind, const void* src, const float* fold, void* dst, bool bf16, int n, int C, int H, int W, int H2, int W2,
               int ld, cudaStream_t stream) {
  if (n <= 0) return TCVN_OK;
  if (C % 8 != 0 || ld % 4 != 0) return fail(TCVN_ERR_UNSUPPORTED, "pool: channel count %d / pitch %d must be multiples of 8 / 4", C, ld);
  if ((long long)n * (H2 > H ? H2 : H) * (W2 > W ? W2 : W) * C >= (1ll << 31) || (long long)n * (H + 2) * (W + 2) * C >= (1ll << 31))
    return fail(TCVN_ERR_UNSUPPORTED, "pool: more than 2^31 elements in one launch (%d images)", n);
  typedef __nv_bfloat16 bf;
  long long total;
  switch (kind) {
    case 0:
      total = (long long)n * H * W * (C / 4);
      if (bf16) stem_pool_fwd_kernel<bf><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), fold, H2, W2, C, static_cast<bf*>(dst), ld, H, W, total);
      else stem_pool_fwd_kernel<float><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), fold, H2, W2, C, static_cast<float*>(dst), ld, H, W, total);
      break;
    case 1:   // gradient buffers of a block are fp32 in both precisions; bf16 = element type of the stem gradient map dst
      total = (long long)n * H2 * W2 * (C / 4);
      if (bf16) stem_pool_bwd_kernel<bf><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), ld, H, W, C, static_cast<bf*>(dst), H2, W2, total);
      else stem_pool_bwd_kernel<float><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), ld, H, W, C, static_cast<float*>(dst), H2, W2, total);
      break;
    case 2:
      total = (long long)n * (H + 2) * (W + 2) * (C / 8);
      if (total >= (1ll << 31)) return fail(TCVN_ERR_UNSUPPORTED, "pool2 backward: too many images in one launch (%d)", n);
      if (bf16) pool2_bwd_kernel<bf><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const bf*>(src), H2, W2, C, static_cast<bf*>(dst), H, W, total);
      else pool2_bwd_kernel<float><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), H2, W2, C, static_cast<float*>(dst), H, W, total);
      break;
    default:
      total = (long long)n * (H + 2) * (W + 2) * C;
      if (bf16) gap_bwd_kernel<bf><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), C, static_cast<bf*>(dst), H, W, total);
      else gap_bwd_kernel<float><<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), C, static_cast<float*>(dst), H, W, total);
      break;
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// BN1 + PReLU1 backward reductions of a dense layer: parts[slab][3][C] (sum g, sum g xhat, sum dA min(y,0))
int stem_pool16_forward(const void* z0_bf16, const float* fold, void* blk_bf16, int n, int C, int H, int W, int Hs, int Ws, int ld,
                        cudaStream_t stream) {
  if (n <= 0) return TCVN_OK;
  if (C % 8 || ld % 8) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: widths must be multiples of 8");
  const long long total = (long long)n * H * W * (C / 8);
  if (total >= (1ll << 31) || (long long)n * Hs * Ws * (C / 8) >= (1ll << 31)) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: too many images in one launch (%d)", n);
  if (H > 65535 || n > 65535) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: grid too large (%d rows, %d images)", H, n);
  const int tpr = W * (C / 8), nb = ceil_div(tpr, 256), bs = (ceil_div(tpr, nb) + 31) & ~31;   // threads per pooled row, evenly split
  stem_pool16_fwd_kernel<<<dim3((unsigned)nb, (unsigned)H, (unsigned)n), bs, 0, stream>>>(static_cast<const __nv_bfloat16*>(z0_bf16), fold, Hs, Ws, C,
                                                                               static_cast<__nv_bfloat16*>(blk_bf16), ld, H, W, total);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int stem_pool16_backward(const void* dblk_bf16, int ld, void* dz_bf16, int n, int C, int H, int W, int Hs, int Ws, cudaStream_t stream) {
  if (n <= 0) return TCVN_OK;
  if (C % 8 || ld % 8) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: widths must be multiples of 8");
  const long long total = (long long)n * Hs * Ws * (C / 8);
  if (total >= (1ll << 31)) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: too many images in one launch (%d)", n);
  if (Hs > 65535 || n > 65535) return fail(TCVN_ERR_UNSUPPORTED, "stem_pool16: grid too large (%d rows, %d images)", Hs, n);
  const int tpr = Ws * (C / 8), nb = ceil_div(tpr, 256), bs = (ceil_div(tpr, nb) + 31) & ~31;
  stem_pool16_bwd_kernel<<<dim3((unsigned)nb, (unsigned)Hs, (unsigned)n), bs, 0, stream>>>(static_cast<const __nv_bfloat16*>(dblk_bf16), ld, H, W, C,
                                                                               static_cast<__nv_bfloat16*>(dz_bf16), Hs, Ws, total);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int bn1_bwd_reduce(const void* X, int ldx, const void* D, int ldd, const float* fold, int fold_stride, int C, long long m_total,
                   int ring_hp, int ring_wp, double* parts, int* n_slabs, cudaStream_t stream) {
  *n_slabs = 0;
  if (m_total <= 0 || C <= 0) return TCVN_OK;
  if (C % 8 || ldx % 8 || ldd % 8) return fail(TCVN_ERR_UNSUPPORTED, "bn1_bwd_reduce: widths must be multiples of 8");
  Bn1FusedDev p{};
  p.X = static_cast<const __nv_bfloat16*>(X); p.ldx = ldx; p.D = static_cast<const __nv_bfloat16*>(D); p.ldd = ldd;
  p.fold = fold; p.fold_stride = fold_stride; p.C = C; p.m_total = m_total; p.Hp = ring_hp; p.Wp = ring_wp;
  p.parts = parts;
  const int cv = C / 8;
  const int tv = cv < 32 ? cv : 32;
  const int rpi = 256 / tv;
  const int gx = ceil_div(cv, tv);
  long long vslabs = ceil_div_ll(m_total, (long long)rpi * 4 * 4);
  const long long cap = (long long)(148 * 2 / gx > 1 ? 148 * 2 / gx : 1);
  if (vslabs > cap) vslabs = cap;
  p.rows_per_slab = (int)ceil_div_ll(m_total, vslabs);
  vslabs = ceil_div_ll(m_total, p.rows_per_slab);
  *n_slabs = (int)vslabs;
  dim3 grid(gx, (unsigned)vslabs);
  bn1_bwd_reduce_vec_kernel<<<grid, 256, 0, stream>>>(p, tv);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

int dropout_typed(void* X, bool bf16, int ld, int col0, int C, long long m_total, uint64_t seed, uint64_t stream_id, float p,
                  cudaStream_t stream) {
  if (p == 0.f || m_total <= 0) return TCVN_OK;
  const unsigned grid = (unsigned)ceil_div_ll(m_total * C, 256);
  if (bf16) dropout_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(X), ld, col0, C, m_total, seed, stream_id, p, seed_offset_ptr());
  else dropout_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(X), ld, col0, C, m_total, seed, stream_id, p, seed_offset_ptr());
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

}  // namespace tcvn

using namespace tcvn;

extern "C" int tcvn_t_gemm(const float* A, int lda, int64_t m_total, int K, int taps, const int32_t* tap_off,
                           const float* W, int N, const float* a_fold, int a_ring_hp, int a_ring_wp, const float* bias,
                           float* out, int ldo, int out_col0, int out_ring_hp, int out_ring_wp, int accumulate,
                           tcvn_stream_t stream) {
  TCVN_CHECK_ARG(A && W && out && taps >= 1 && taps <= 9 && (taps == 1 || tap_off), "t_gemm: bad arguments");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.m_total = m_total; g.K = K; g.taps = taps;
  for (int t = 0; t < taps; ++t) g.tap_off[t] = tap_off ? tap_off[t] : 0;
  g.W = W; g.N = N;
  if (a_fold) { g.a_scale = a_fold; g.a_shift = a_fold + K; g.a_alpha = a_fold + 2 * K; }
  g.o_shift = bias;
  g.out = out; g.ldo = ldo; g.out_col0 = out_col0; g.ring_Hp = out_ring_hp; g.ring_Wp = out_ring_wp;
  g.a_is_f32 = true; g.out_is_f32 = true; g.a_ring_Hp = a_ring_hp; g.a_ring_Wp = a_ring_wp; g.accumulate = accumulate != 0;
  return launch_simt_gemm(g, stream);
}

namespace tcvn {
int wgrad_typed(const void* A, bool a_bf16, int lda, long long m_total, int K, int taps, const int* tap_off, const float* a_scale,
                const float* a_shift, const float* a_alpha, int a_ring_hp, int a_ring_wp, const void* G, bool g_bf16, int ldg,
                int g_col0, int N, int g_ring_hp, int g_ring_wp, float* dW, cudaStream_t stream, float* parts,
                size_t parts_floats) {
  if (m_total <= 0) return TCVN_OK;
  WgradDev g{};
  g.A = A; g.lda = lda; g.m_total = m_total; g.K = K; g.taps = taps;
  for (int t = 0; t < taps; ++t) g.tap_off[t] = tap_off ? tap_off[t] : 0;
  g.a_scale = a_scale; g.a_shift = a_shift; g.a_alpha = a_alpha;
  g.a_ring_Hp = a_ring_hp; g.a_ring_Wp = a_ring_wp;
  g.G = G; g.ldg = ldg; g.g_col0 = g_col0; g.N = N; g.g_ring_Hp = g_ring_hp; g.g_ring_Wp = g_ring_wp;
  g.dW = dW;
  int slabs = slabs_for(m_total, &g.rows_per_slab);
  // Bit-reproducible whenever possible: with a scratch buffer every row slab stores its partial result in its own slot and
  // the slots are added in a fixed order; without one, up to 4096 rows are walked by a single CTA per result tile (one
  // writer per element).  Only larger inputs without scratch (the fp32 parity walk) add with float atomics.
  const size_t n_out = (size_t)taps * K * N;
  bool slotted = false;
  if (slabs > 1 && parts != nullptr && n_out % 4 == 0 && n_out * slabs <= parts_floats) slotted = true;
  else if (m_total <= 4096) { slabs = 1; g.rows_per_slab = (int)m_total; }
  g.parts = slotted ? parts : nullptr;
  dim3 grid(ceil_div(K, kWgK), ceil_div(N, kWgN) * taps, slabs);
  typedef __nv_bfloat16 bf;
  if (!a_bf16 && !g_bf16) wgrad_kernel<float, float><<<grid, 256, 0, stream>>>(g);
  else if (a_bf16 && g_bf16) wgrad_kernel<bf, bf><<<grid, 256, 0, stream>>>(g);
  else if (a_bf16) wgrad_kernel<bf, float><<<grid, 256, 0, stream>>>(g);
  else wgrad_kernel<float, bf><<<grid, 256, 0, stream>>>(g);
  TCVN_LAUNCH_CHECK();
  if (slotted) TCVN_TRY(reduce_parts(parts, slabs, (long long)n_out, dW, true, stream));
  return TCVN_OK;
}

int wgrad_f32(const float* A, int lda, long long m_total, int K, const float* G, int ldg, int N, float* dW, float* parts,
              size_t parts_floats, cudaStream_t stream) {
  return wgrad_typed(A, false, lda, m_total, K, 1, nullptr, nullptr, nullptr, nullptr, 0, 0, G, false, ldg, 0, N, 0, 0, dW, stream,
                     parts, parts_floats);
}
}  // namespace tcvn

extern "C" int tcvn_t_wgrad(const float* A, int lda, int64_t m_total, int K, int taps, const int32_t* tap_off,
                            const float* a_fold, int a_ring_hp, int a_ring_wp, const float* G, int ldg, int g_col0, int N,
                            int g_ring_hp, int g_ring_wp, float* dW, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(A && G && dW && taps >= 1 && taps <= 9 && (taps == 1 || tap_off), "t_wgrad: bad arguments");
  return wgrad_typed(A, false, lda, m_total, K, taps, tap_off, a_fold, a_fold ? a_fold + K : nullptr,
                     a_fold ? a_fold + 2 * K : nullptr, a_ring_hp, a_ring_wp, G, false, ldg, g_col0, N, g_ring_hp, g_ring_wp, dW,
                     stream);
}

// mode 0: sums[2][C] = (sum x, sum x^2); mode 1: sums[3][C] BN+PReLU backward reductions; mode 2: sums[1][C] = sum x.
// sums is overwritten (one slab of rows per column block: bit-reproducible, meant for tests and small inputs).
extern "C" int tcvn_t_colsums(int mode, const float* X, int ldx, int xcol0, const float* D, int ldd, int dcol0,
                              const float* fold, int C, int64_t m_total, int ring_hp, int ring_wp, double* sums,
                              tcvn_stream_t stream) {
  TCVN_CHECK_ARG(X && sums && mode >= 0 && mode <= 2 && (mode != 1 || (D && fold)), "t_colsums: bad arguments");
  return colsums_into(mode, X, ldx, xcol0, D, ldd, dcol0, fold, C, m_total, ring_hp, ring_wp, sums, C, nullptr, stream);
}

extern "C" int tcvn_t_bn_finalize(const double* sums, int C, double count, const float* gamma, const float* beta,
                                  const float* alpha, float eps, float momentum, float* running_mean, float* running_var,
                                  float* fold, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(sums && gamma && beta && fold && count > 0, "t_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, C, count, gamma, beta, alpha, eps, momentum, running_mean,
                                                           running_var, fold);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_t_bnact_bwd_apply(const float* D, int ldd, int dcol0, const float* X, int ldx, int xcol0,
                                      const float* fold, const double* sums, int C, double count, float* dX, int lddx,
                                      int dxcol0, int accumulate, int64_t m_total, int ring_hp, int ring_wp, float* dgamma,
                                      float* dbeta, float* dalpha, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(D && X && fold && sums && count > 0, "t_bnact_bwd_apply: bad arguments");
  if (dX) TCVN_TRY(bnact_bwd_apply_typed(D, false, ldd, dcol0, X, false, ldx, xcol0, fold, C, sums, C, count, dX, false, lddx, dxcol0,
                                         accumulate != 0, m_total, ring_hp, ring_wp, stream));
  if (dgamma || dbeta || dalpha) {
    bnact_param_grads_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, C, dgamma, dbeta, dalpha);
    TCVN_LAUNCH_CHECK();
  }
  return TCVN_OK;
}

extern "C" int tcvn_t_add_colsums(const double* sums, int C, float* dst, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(sums && dst, "t_add_colsums: null pointer");
  add_cols_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, C, dst);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_t_bnact_fwd(const float* X, int ldx, int xcol0, const float* fold, int C, int64_t m_total, int ring_hp,
                                int ring_wp, float* out, int ldo, int ocol0, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(X && fold && out, "t_bnact_fwd: null pointer");
  return bnact_fwd_typed(X, false, ldx, xcol0, fold, C, C, m_total, ring_hp, ring_wp, out, false, ldo, ocol0, stream);
}

// kind 0: stem pool fwd (src = z0 [n,Hs,Ws,C], fold; dst = ringed blk, ld);  kind 1: stem pool bwd (src = d blk, ld; dst = dA
// [n,Hs,Ws,C]);  kind 2: pool2 bwd (src = dP ringed [n,H2+2,W2+2,C]; dst = dA ringed [n,H+2,W+2,C]);
// kind 3: gap bwd (src = dGap [n,C]; dst = dA ringed [n,H+2,W+2,C])
extern "C" int tcvn_t_pool(int kind, const float* src, const float* fold, float* dst, int n, int C, int H, int W, int H2,
                           int W2, int ld, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(src && dst && kind >= 0 && kind <= 3, "t_pool: bad arguments");
  return pool_typed(kind, src, fold, dst, false, n, C, H, W, H2, W2, ld, stream);
}

extern "C" int tcvn_t_dropout(float* X, int ld, int col0, int C, int64_t m_total, uint64_t seed, uint64_t stream_id, float p,
                              tcvn_stream_t stream) {
  TCVN_CHECK_ARG(X && p >= 0.f && p < 1.f, "t_dropout: bad arguments");
  return dropout_typed(X, false, ld, col0, C, m_total, seed, stream_id, p, stream);
}

// forward (dz == NULL): z[n,Hs,Ws,C] = bias + conv7x7s2p3(pixels);  backward (dz != NULL): dw[cin*49][C] += x (*) dz
namespace tcvn {
// conv0 forward (dz == nullptr: z = bias + conv) or weight gradient (dw += pixels (x) dz); dz_bf16 = element type of dz
int stem_conv_typed(const float* pixels, int n, int cin, int H, int W, const float* w, const float* bias, int C, float* z,
                    const void* dz, bool dz_bf16, float* dw, cudaStream_t stream) {
  TCVN_CHECK_ARG(pixels && w && ((dz == nullptr && z && bias) || (dz && dw)), "t_stem_conv: bad arguments");
  if (n <= 0) return TCVN_OK;
  const int Hs = (H + 6 - 7) / 2 + 1, Ws = (W + 6 - 7) / 2 + 1;
  if (dz == nullptr) {
    const long long outs4 = (long long)n * Hs * Ws * C / 4;
    if (outs4 >= (1ll << 30)) return fail(TCVN_ERR_UNSUPPORTED, "t_stem_conv: %d images exceed 2^32 stem outputs in one launch", n);
    stem_fill_bias_kernel<<<(unsigned)ceil_div_ll(outs4, 256), 256, 0, stream>>>(z, bias, C, outs4);
    TCVN_LAUNCH_CHECK();
  }
  const long long total = (long long)n * cin * H * W;
  if (total >= (1ll << 32) - (1ll << 22)) return fail(TCVN_ERR_UNSUPPORTED, "t_stem_conv: %d images exceed 2^32 pixel values in one launch", n);
  if (cin * 49 * C > kStemW) return fail(TCVN_ERR_UNSUPPORTED, "t_stem_conv: filter bank larger than %d entries", kStemW);
  long long want = ceil_div_ll(total, 256 * 8);
  const int grid = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
  if (dz == nullptr)
    stem_conv_scatter_kernel<false><<<grid, 256, 0, stream>>>(pixels, cin, H, W, Hs, Ws, w, C, z, static_cast<const float*>(nullptr), dw, (unsigned)total);
  else if (dz_bf16)
    stem_conv_scatter_kernel<true, __nv_bfloat16><<<grid, 256, 0, stream>>>(pixels, cin, H, W, Hs, Ws, w, C, z,
                                                                            static_cast<const __nv_bfloat16*>(dz), dw, (unsigned)total);
  else
    stem_conv_scatter_kernel<true><<<grid, 256, 0, stream>>>(pixels, cin, H, W, Hs, Ws, w, C, z, static_cast<const float*>(dz), dw,
                                                             (unsigned)total);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
}  // namespace tcvn

extern "C" int tcvn_t_stem_conv(const float* pixels, int n, int cin, int H, int W, const float* w, const float* bias, int C,
                                float* z, const float* dz, float* dw, tcvn_stream_t stream) {
  return tcvn::stem_conv_typed(pixels, n, cin, H, W, w, bias, C, z, dz, false, dw, stream);
}

// ------------------------------------------------------------------------------------------------
// fused AdamW over a flat fp32 buffer (torch.optim.AdamW semantics, trainers/neutrino_base.py:109-130), with the
// global-norm clip factor of Lightning's gradient_clip_val (train.py:140) read from device memory: no host sync.
// ------------------------------------------------------------------------------------------------
namespace tcvn {

constexpr int kMaxOptGroups = 4;
constexpr int kSumsqBlocksMax = 148 * 8;

// sum of squares, stage 1: one partial sum per block in its own slot (no atomics: the clip factor is bit-reproducible)
__global__ void __launch_bounds__(256) sumsq_parts_kernel(const float* __restrict__ x, long long n, double* __restrict__ parts) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    acc += v * v;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) parts[blockIdx.x] = red[0];
}

struct AdamGroup { float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt; double beta1d, beta2d; };
struct AdamArgs {
  float* p; const float* g; float* m; float* v; long long n;
  const uint8_t* select;          // per element: 0 = not optimised, k = group k-1
  int n_groups; AdamGroup grp[kMaxOptGroups];
  const double* gnorm_parts; int n_parts;   // stage-1 partial sums of the squared gradient norm (nullptr: no clipping)
  float max_norm, grad_mul;
  double* gnorm_out;              // optional: the total norm squared, written by block 0
  // device-resident overrides (a captured CUDA graph of the step bakes the by-value arguments in):
  const long long* step_dev;      // optimizer step count (>= 1): bias corrections are then computed on the device
  const float* lr_dev;            // [n_groups] learning rates (the LR scheduler's per-step values)
};

// every parameter group in ONE pass over the arenas; each block first adds the norm partials in the same fixed order
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamArgs a) {
  __shared__ AdamGroup sgrp[kMaxOptGroups];
  if (threadIdx.x < a.n_groups) {
    AdamGroup h = a.grp[threadIdx.x];
    if (a.lr_dev) h.lr = a.lr_dev[threadIdx.x];
    if (a.step_dev) {
      const double st = (double)*a.step_dev;
      h.bc1 = (float)(1.0 - pow(h.beta1d, st));          // the host path's expression, in double, on the device
      h.bc2_sqrt = (float)sqrt(1.0 - pow(h.beta2d, st));
    }
    sgrp[threadIdx.x] = h;
  }
  __syncthreads();
  float clip = a.grad_mul;
  if (a.gnorm_parts != nullptr && a.max_norm > 0.f) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < a.n_parts; i += 256) acc += a.gnorm_parts[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    const double total = red[0];
    if (a.gnorm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.gnorm_out = total;
    const float norm = (float)sqrt(total) * a.grad_mul;
    const float coef = a.max_norm / (norm + 1e-6f);
    if (coef < 1.f) clip *= coef;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const int gi = a.select != nullptr ? (int)a.select[i] : 1;
    if (gi == 0 || gi > a.n_groups) continue;
    const AdamGroup& h = sgrp[gi - 1];
    const float grad = a.g[i] * clip;
    float w = a.p[i] * (1.f - h.lr * h.weight_decay);
    const float mi = h.beta1 * a.m[i] + (1.f - h.beta1) * grad;
    const float vi = h.beta2 * a.v[i] + (1.f - h.beta2) * grad * grad;
    a.m[i] = mi;
    a.v[i] = vi;
    const float denom = sqrtf(vi) / h.bc2_sqrt + h.eps;
    w -= (h.lr / h.bc1) * (mi / denom);
    a.p[i] = w;
  }
}

}  // namespace tcvn

extern "C" size_t tcvn_adamw_workspace_bytes(void) { return (size_t)(tcvn::kSumsqBlocksMax + 1) * sizeof(double); }

// One optimizer step over the flat arenas: every parameter group (select[i] = group + 1, 0 = element not optimised) in one
// launch, preceded - when max_norm > 0 - by the global gradient-norm reduction for Lightning's gradient_clip_val.
// workspace: tcvn_adamw_workspace_bytes(); its last double receives the squared gradient norm.
extern "C" int tcvn_adamw_fused(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const uint8_t* select, int n_groups, const double* lr, const double* beta1, const double* beta2,
                                const double* eps, const double* weight_decay, const int64_t* step, float max_norm,
                                float grad_mul, void* workspace, size_t workspace_bytes, const int64_t* step_dev,
                                const float* lr_dev, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 0 && lr && beta1 && beta2 && eps && weight_decay && step,
                 "adamw_fused: null pointer");
  TCVN_CHECK_ARG(n_groups >= 1 && n_groups <= tcvn::kMaxOptGroups, "adamw_fused: 1..%d parameter groups (got %d)",
                 tcvn::kMaxOptGroups, n_groups);
  TCVN_CHECK_ARG(n_groups == 1 || select, "adamw_fused: several groups need the per-element selector");
  if (n == 0) return TCVN_OK;
  tcvn::AdamArgs a{};
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n = n; a.select = select; a.n_groups = n_groups;
  for (int k = 0; k < n_groups; ++k) {
    TCVN_CHECK_ARG(step[k] >= 1, "adamw_fused: step must be >= 1");
    // bias corrections in double on the host, like torch.optim.AdamW's Python scalars
    a.grp[k].lr = (float)lr[k]; a.grp[k].beta1 = (float)beta1[k]; a.grp[k].beta2 = (float)beta2[k];
    a.grp[k].eps = (float)eps[k]; a.grp[k].weight_decay = (float)weight_decay[k];
    a.grp[k].beta1d = beta1[k]; a.grp[k].beta2d = beta2[k];
    a.grp[k].bc1 = (float)(1.0 - pow(beta1[k], (double)step[k]));
    a.grp[k].bc2_sqrt = (float)sqrt(1.0 - pow(beta2[k], (double)step[k]));
  }
  a.max_norm = max_norm; a.grad_mul = grad_mul;
  a.step_dev = reinterpret_cast<const long long*>(step_dev); a.lr_dev = lr_dev;
  if (max_norm > 0.f) {
    TCVN_CHECK_ARG(workspace && workspace_bytes >= tcvn_adamw_workspace_bytes(), "adamw_fused: workspace too small for the norm");
    double* parts = static_cast<double*>(workspace);
    int blocks = (int)tcvn::ceil_div_ll(n, 256 * 8);
    if (blocks > tcvn::kSumsqBlocksMax) blocks = tcvn::kSumsqBlocksMax;
    tcvn::sumsq_parts_kernel<<<blocks, 256, 0, stream>>>(grads, n, parts);
    TCVN_LAUNCH_CHECK();
    a.gnorm_parts = parts; a.n_parts = blocks; a.gnorm_out = parts + tcvn::kSumsqBlocksMax;
  }
  int blocks = (int)tcvn::ceil_div_ll(n, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tcvn::adamw_multi_kernel<<<blocks, 256, 0, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// fused BN + PReLU + AvgPool2d(2,2) (transition front half) and BN + PReLU + global average (tail), fp32, with a
// batch-statistics fold: the same kernels the inference path uses
extern "C" int tcvn_t_act_pool2(const float* blk, int n, int H, int W, int ld, int C, const float* fold, float* out, int H2,
                                int W2, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(blk && fold && out, "t_act_pool2: null pointer");
  return tcvn::launch_act_pool2(blk, n, H, W, ld, C, fold, fold + C, fold + 2 * C, out, H2, W2, true, stream);
}

extern "C" int tcvn_t_act_gap(const float* blk, int n, int H, int W, int ld, int C, const float* fold, float* gap,
                              tcvn_stream_t stream) {
  TCVN_CHECK_ARG(blk && fold && gap, "t_act_gap: null pointer");
  return tcvn::launch_act_gap(blk, n, H, W, ld, C, fold, fold + C, fold + 2 * C, gap, true, stream);
}
