// Training primitives of the sequence part (token assembly, post-norm transformer encoder, heads), fp32.
// Token rows are a plain matrix X[R = B*S, D], row = event*S + slot (slot 0 = event token).
// Reference: transformercvn/network/layers/prong_custom_bert_encoder.py:45-75 (torch TransformerEncoderLayer,
// post-norm, gelu(erf), key padding mask), networks/neutrino_full_base_network.py:99-125,
// layers/packed_data.py:59-76.
#include "kernels.h"

namespace tcvn {

// out = LayerNorm(a + b) * gamma + beta ; saves the pre-norm sum and (mean, rstd) per row.  One warp per row, D <= 128.
__global__ void ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int D, int R,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                              float* __restrict__ pre, float* __restrict__ stats, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float v[4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < D ? a[(size_t)row * D + j] + (b ? b[(size_t)row * D + j] : 0.f) : 0.f;
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float d = lane + 32 * i < D ? v[i] - mean : 0.f;
    var += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    if (j < D) {
      pre[(size_t)row * D + j] = v[i];
      out[(size_t)row * D + j] = (v[i] - mean) * rstd * gamma[j] + beta[j];
    }
  }
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.  The parameter gradients are a column reduction over
// the rows and are taken by ln_param_grads_kernel in a fixed order (float atomics here made the step irreproducible).
__global__ void ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, const float* __restrict__ stats,
                              const float* __restrict__ gamma, int D, int R, float* __restrict__ dx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float g[4], xh[4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    g[i] = 0.f; xh[i] = 0.f;
    if (j < D) {
      const float d = dy[(size_t)row * D + j];
      xh[i] = (pre[(size_t)row * D + j] - mean) * rstd;
      g[i] = d * gamma[j];
      s1 += g[i];
      s2 = fmaf(g[i], xh[i], s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 /= (float)D; s2 /= (float)D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    if (j < D) dx[(size_t)row * D + j] = rstd * (g[i] - s1 - xh[i] * s2);
  }
}

// dgamma[j] += sum_r dy[r,j] * xhat[r,j] ; dbeta[j] += sum_r dy[r,j].  A block owns 32 columns, its 32 warps take every
// 32nd row, the 32 partial sums are added in order: one writer per element, fixed summation order.
__global__ void __launch_bounds__(1024) ln_param_grads_kernel(const float* __restrict__ dy, const float* __restrict__ pre,
                                                             const float* __restrict__ stats, int D, int R, float* dgamma,
                                                             float* dbeta) {
  __shared__ float red[2][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float sg = 0.f, sb = 0.f;
  if (j < D)
    for (int r = w; r < R; r += 32) {
      const float d = dy[(size_t)r * D + j];
      sg = fmaf(d, (pre[(size_t)r * D + j] - stats[2 * r]) * stats[2 * r + 1], sg);
      sb += d;
    }
  red[0][w][lane] = sg;
  red[1][w][lane] = sb;
  __syncthreads();
  if (w == 0 && j < D) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) { a += red[0][k][lane]; b += red[1][k][lane]; }
    dgamma[j] += a;
    dbeta[j] += b;
  }
}

__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long stream_id, unsigned long long idx,
                                          float p) {
  unsigned long long z = seed * 0x100000001b3ull + stream_id * 0x9e3779b97f4a7c15ull + idx;
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  const uint32_t r = (uint32_t)((z ^ (z >> 31)) >> 32);
  return (r >> 8) * (1.0f / 16777216.0f) >= p;
}

// multi-head attention over one event and one head per CTA; S <= 32 tokens, one thread per query.
// qkv[R][3D]; P[b][h][S][S] = softmax probabilities (saved BEFORE dropout; the mask is re-derived from the seed); ctx[R][D]
__global__ void attn_fwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int S, int heads, int D,
                                float* __restrict__ P, float* __restrict__ ctx, float p_drop, unsigned long long seed0,
                                unsigned long long stream_id, const unsigned long long* seed_off) {
  const unsigned long long seed = seed_with_offset(seed0, seed_off);
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int s = threadIdx.x;
  if (s >= S) return;
  const int dh = D / heads;
  const float scale = rsqrtf((float)dh);
  const float* q = qkv + ((size_t)(b * S + s)) * 3 * D + h * dh;
  float sc[32];
  float mx = -INFINITY;
  for (int t = 0; t < S; ++t) {
    float dot = -INFINITY;
    if (mask[b * S + t]) {
      const float* k = qkv + ((size_t)(b * S + t)) * 3 * D + D + h * dh;
      dot = 0.f;
      for (int j = 0; j < dh; ++j) dot = fmaf(q[j] * scale, k[j], dot);
    }
    sc[t] = dot;
    mx = fmaxf(mx, dot);
  }
  float den = 0.f;
  for (int t = 0; t < S; ++t) {
    const float e = mask[b * S + t] ? expf(sc[t] - mx) : 0.f;
    sc[t] = e;
    den += e;
  }
  const float inv = 1.f / den;
  float* prow = P + (((size_t)b * heads + h) * S + s) * S;
  for (int t = 0; t < S; ++t) {
    float pv = sc[t] * inv;
    prow[t] = pv;
    if (p_drop > 0.f)
      pv = drop_keep(seed, stream_id, (((size_t)b * heads + h) * S + s) * S + t, p_drop) ? pv / (1.f - p_drop) : 0.f;
    sc[t] = pv;
  }
  for (int j = 0; j < dh; ++j) {
    float o = 0.f;
    for (int t = 0; t < S; ++t) o = fmaf(sc[t], qkv[((size_t)(b * S + t)) * 3 * D + 2 * D + h * dh + j], o);
    ctx[((size_t)(b * S + s)) * D + h * dh + j] = o;
  }
}

// backward of the above, one (event, head) per warp.  Pass 1: query s derives its rows of dP (through the dropout mask,
// re-derived from (seed, stream, index)) and dS into shared memory and its own dQ.  Pass 2: key t gathers dK[t] and
// dV[t] over the queries in order - every element of dqkv has one writer and a fixed summation order (no atomics, no
// memset of dqkv).
__global__ void __launch_bounds__(32) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                      const float* __restrict__ dctx, const uint8_t* __restrict__ mask, int S,
                                                      int heads, int D, float p_drop, unsigned long long seed0,
                                                      unsigned long long stream_id, float* __restrict__ dqkv,
                                                      const unsigned long long* seed_off) {
  const unsigned long long seed = seed_with_offset(seed0, seed_off);
  __shared__ float pv_s[32][33];   // [query][key]: what multiplied V in the forward pass
  __shared__ float ds_s[32][33];   // [query][key]: gradient of the scaled scores
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int s = threadIdx.x;
  const int dh = D / heads;
  const float scale = rsqrtf((float)dh);
  if (s < S) {
    const float* prow = P + (((size_t)b * heads + h) * S + s) * S;
    const float* dc = dctx + ((size_t)(b * S + s)) * D + h * dh;
    float dP[32];
    float dot = 0.f;  // sum_t dPsoft[t] * Psoft[t]
    for (int t = 0; t < S; ++t) {
      const float psoft = prow[t];
      float keep = 1.f;
      if (p_drop > 0.f)
        keep = drop_keep(seed, stream_id, (((size_t)b * heads + h) * S + s) * S + t, p_drop) ? 1.f / (1.f - p_drop) : 0.f;
      float d = 0.f;
      const float* v = qkv + ((size_t)(b * S + t)) * 3 * D + 2 * D + h * dh;
      for (int j = 0; j < dh; ++j) d = fmaf(dc[j], v[j], d);
      pv_s[s][t] = psoft * keep;
      dP[t] = d * keep;
      dot = fmaf(dP[t], psoft, dot);
    }
    float dq[16];
    for (int j = 0; j < dh; ++j) dq[j] = 0.f;
    for (int t = 0; t < S; ++t) {
      float dS = 0.f;
      if (mask[b * S + t]) {
        dS = prow[t] * (dP[t] - dot) * scale;
        const float* k = qkv + ((size_t)(b * S + t)) * 3 * D + D + h * dh;
        for (int j = 0; j < dh; ++j) dq[j] = fmaf(dS, k[j], dq[j]);
      }
      ds_s[s][t] = dS;
    }
    for (int j = 0; j < dh; ++j) dqkv[((size_t)(b * S + s)) * 3 * D + h * dh + j] = dq[j];
  }
  __syncwarp();
  if (s < S) {
    const int t = s;   // this thread's key / value row
    float dk[16], dv[16];
    for (int j = 0; j < dh; ++j) { dk[j] = 0.f; dv[j] = 0.f; }
    for (int q = 0; q < S; ++q) {
      const float pv = pv_s[q][t], dS = ds_s[q][t];
      const float* dc = dctx + ((size_t)(b * S + q)) * D + h * dh;
      const float* qq = qkv + ((size_t)(b * S + q)) * 3 * D + h * dh;
      for (int j = 0; j < dh; ++j) {
        dv[j] = fmaf(pv, dc[j], dv[j]);
        dk[j] = fmaf(dS, qq[j], dk[j]);
      }
    }
    for (int j = 0; j < dh; ++j) {
      dqkv[((size_t)(b * S + t)) * 3 * D + D + h * dh + j] = dk[j];
      dqkv[((size_t)(b * S + t)) * 3 * D + 2 * D + h * dh + j] = dv[j];
    }
  }
}

// elementwise helpers: kind 0 gelu fwd (out = gelu(x)); 1 gelu bwd (out = d * gelu'(x)); 2 out = a + b; 3 rows *= mask
__global__ void eltwise_kernel(int kind, const float* __restrict__ a, const float* __restrict__ b,
                               const uint8_t* __restrict__ mask, int C, long long total, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (kind == 0) {
    const float x = a[i];
    out[i] = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  } else if (kind == 1) {
    const float x = a[i];
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    out[i] = b[i] * (cdf + x * pdf);
  } else if (kind == 2) {
    out[i] = a[i] + b[i];
  } else {
    out[i] = mask[i / C] ? a[i] : 0.f;
  }
}

// token assembly.  dir 0: build Xtok[(B+T)][in_dim] from embeddings; dir 1: scatter its gradient back
// (dev_emb, dpr_emb overwritten; dpos accumulated by the caller with a column sum of the position columns)
__global__ void tok_gather_kernel(int dir, float* ev_emb, float* pr_emb, const float* __restrict__ pos, int B, int T,
                                  int pixel_dim, int feature_dim, int position_dim, float* xtok) {
  const int in_dim = pixel_dim + feature_dim + position_dim;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(B + T) * in_dim) return;
  const int k = (int)(i % in_dim);
  const int r = (int)(i / in_dim);
  const int ev_dim = pixel_dim + feature_dim;
  if (dir == 0) {
    float v;
    if (r < B) v = k < ev_dim ? ev_emb[(size_t)r * ev_dim + k] : pos[k - ev_dim];
    else if (k < feature_dim) v = 0.f;
    else if (k < ev_dim) v = pr_emb[(size_t)(r - B) * pixel_dim + k - feature_dim];
    else v = pos[k - ev_dim];
    xtok[i] = v;
  } else {
    if (r < B) { if (k < ev_dim) ev_emb[(size_t)r * ev_dim + k] = xtok[i]; }
    else if (k >= feature_dim && k < ev_dim) pr_emb[(size_t)(r - B) * pixel_dim + k - feature_dim] = xtok[i];
  }
}

// rows[(B+T)][D] <-> X[B*S][D].  dir 0: scatter rows into the padded sequence (padded slots zero); dir 1: gather grads.
__global__ void tok_scatter_kernel(int dir, float* rows, float* X, const uint8_t* __restrict__ prong_mask,
                                   const int* __restrict__ prong_offset, int B, int L, int D) {
  const int S = 1 + L;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * S * D) return;
  const int j = (int)(i % D);
  const int s = (int)((i / D) % S);
  const int b = (int)(i / ((long long)D * S));
  int src = -1;
  if (s == 0) src = b;
  else if (prong_mask[(size_t)b * L + s - 1]) {
    int rank = 0;
    for (int t = 0; t < s - 1; ++t) rank += prong_mask[(size_t)b * L + t] != 0;
    src = B + prong_offset[b] + rank;
  }
  if (dir == 0) X[i] = src >= 0 ? rows[(size_t)src * D + j] : 0.f;
  else if (src >= 0) rows[(size_t)src * D + j] = X[i];
}

// prong-head rows: out[(l*B + b)] = X[b*S + 1 + l]  (dir 0) / X[b*S + 1 + l] += out[...] (dir 1); the order (l, b)
// is the reference's (prong_target_decoder.py:35-39 reshapes (L,B,D) -> (L*B,D)); BN1d statistics do not depend on it
__global__ void head_rows_kernel(int dir, float* X, float* rows, int B, int L, int D) {
  const int S = 1 + L;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * L * D) return;
  const int j = (int)(i % D);
  const int b = (int)((i / D) % B);
  const int l = (int)(i / ((long long)D * B));
  float* x = X + ((size_t)b * S + 1 + l) * D + j;
  if (dir == 0) rows[i] = *x; else *x += rows[i];
}

__global__ void prong_offsets_train_kernel(const uint8_t* __restrict__ prong_mask, int B, int L, int* __restrict__ offsets) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int run = 0;
  for (int b = 0; b < B; ++b) {
    offsets[b] = run;
    for (int l = 0; l < L; ++l) run += prong_mask[(size_t)b * L + l] != 0;
  }
  offsets[B] = run;
}

}  // namespace tcvn

using namespace tcvn;

extern "C" int tcvn_t_layernorm(int dir, const float* a, const float* b, int D, int R, const float* gamma, const float* beta,
                                float eps, float* pre, float* stats, float* out, float* dgamma, float* dbeta,
                                tcvn_stream_t stream) {
  TCVN_CHECK_ARG(a && gamma && pre && stats && out && D <= 128, "t_layernorm: bad arguments");
  if (R <= 0) return TCVN_OK;
  if (dir == 0) ln_fwd_kernel<<<ceil_div(R, 8), 256, 0, stream>>>(a, b, D, R, gamma, beta, eps, pre, stats, out);
  else {
    ln_bwd_kernel<<<ceil_div(R, 8), 256, 0, stream>>>(a, pre, stats, gamma, D, R, out);
    TCVN_LAUNCH_CHECK();
    if (dgamma && dbeta) ln_param_grads_kernel<<<ceil_div(D, 32), 1024, 0, stream>>>(a, pre, stats, D, R, dgamma, dbeta);
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_t_attention(int dir, const float* qkv, const uint8_t* mask, int B, int S, int heads, int D, float* P,
                                float* ctx_or_dctx, float* dqkv, float p_drop, uint64_t seed, uint64_t stream_id,
                                tcvn_stream_t stream) {
  TCVN_CHECK_ARG(qkv && mask && P && ctx_or_dctx && S <= 32 && D / heads <= 16, "t_attention: bad arguments");
  if (B <= 0) return TCVN_OK;
  if (dir == 0) attn_fwd_kernel<<<B * heads, 32, 0, stream>>>(qkv, mask, S, heads, D, P, ctx_or_dctx, p_drop, seed, stream_id, seed_offset_ptr());
  else {
    TCVN_CHECK_ARG(dqkv, "t_attention: dqkv missing");
    attn_bwd_kernel<<<B * heads, 32, 0, stream>>>(qkv, P, ctx_or_dctx, mask, S, heads, D, p_drop, seed, stream_id, dqkv, seed_offset_ptr());
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_t_eltwise(int kind, const float* a, const float* b, const uint8_t* mask, int C, int64_t total, float* out,
                              tcvn_stream_t stream) {
  TCVN_CHECK_ARG(a && out && kind >= 0 && kind <= 3, "t_eltwise: bad arguments");
  if (total <= 0) return TCVN_OK;
  eltwise_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(kind, a, b, mask, C, total, out);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_t_tokens(int op, int dir, float* a, float* b, const float* pos, const uint8_t* prong_mask, int* offsets,
                             int B, int T, int L, int D, int pixel_dim, int feature_dim, int position_dim, float* x,
                             tcvn_stream_t stream) {
  // op 0: offsets from mask; op 1: gather token inputs (a = ev_emb, b = pr_emb, x = Xtok); op 2: scatter rows (a) <-> X (x);
  // op 3: prong-head rows (a = X, b = rows)
  if (op == 0) {
    prong_offsets_train_kernel<<<1, 32, 0, stream>>>(prong_mask, B, L, offsets);
  } else if (op == 1) {
    const long long total = (long long)(B + T) * (pixel_dim + feature_dim + position_dim);
    if (total > 0)
      tok_gather_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(dir, a, b, pos, B, T, pixel_dim, feature_dim,
                                                                                position_dim, x);
  } else if (op == 2) {
    const long long total = (long long)B * (1 + L) * D;
    if (total > 0) tok_scatter_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(dir, a, x, prong_mask, offsets, B, L, D);
  } else if (op == 3) {
    const long long total = (long long)B * L * D;
    if (total > 0) head_rows_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(dir, a, b, B, L, D);
  } else {
    return fail(TCVN_ERR_ARG, "t_tokens: unknown op %d", op);
  }
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
