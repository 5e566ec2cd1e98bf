// Token assembly + 6-layer post-norm transformer encoder + event / prong heads, eval mode, fp32.
// One CTA per event keeps the whole <=32-token sequence in shared memory for all stages; the
// reference spends ~100 ATen launches here, this is one.
//
// Reference arithmetic:
//   tokens   transformercvn/network/networks/neutrino_full_base_network.py:99-125
//            (prong rows use event_position_embedding, :107; smart features are zeros,
//             layers/prong_feature_embedding.py:75-76; LinearBlock = Linear(no bias)->BN1d->PReLU, :7-33)
//   encoder  transformercvn/network/layers/prong_custom_bert_encoder.py:45-75 around torch's
//            TransformerEncoderLayer (post-norm, gelu(erf), key padding mask, 1/sqrt(d_head) scaling)
//   heads    layers/prong_decoder.py:13-16 ; layers/prong_target_decoder.py:19-41 over ALL slots
#include "kernels.h"

namespace tcvn {

constexpr int kSeqMax = 32;      // 1 + max prongs (the dataset caps prongs at 20)
constexpr int kSeqThreads = 256;

struct SeqPlan {
  tcvn_seq_desc d;
  int in_dim;  // feature + pixel + position
  size_t p_pos, p_cw, p_c_scale, p_c_shift, p_c_alpha;
  struct Layer { size_t wqkv, bqkv, wo, bo, w1, b1, w2, b2, ln1w, ln1b, ln2w, ln2b; } layer[16];
  size_t p_ew, p_eb;
  struct Dec { size_t w, scale, shift, alpha; int cin, cout; } dec[TCVN_MAX_DECODER_LAYERS];
  size_t p_ow, p_ob;
  size_t packed_bytes;
  static bool build(const tcvn_seq_desc& d, SeqPlan* P) {
    if (d.layers < 0 || d.layers > 16 || d.hidden < 1 || d.hidden % d.heads || d.ffn > 3 * d.hidden) return false;
    if (d.num_decoder_layers < 0 || d.num_decoder_layers > TCVN_MAX_DECODER_LAYERS) return false;
    if (d.hidden > 128 || d.hidden % 32) return false;  // register tiling of the LayerNorm / attention code
    P->d = d;
    P->in_dim = d.feature_dim + d.pixel_dim + d.position_dim;
    size_t p = 0;
    auto take = [&](size_t floats) { size_t o = p; p += (floats * 4 + 255) / 256 * 256; return o; };
    const int D = d.hidden;
    P->p_pos = take(d.position_dim);
    P->p_cw = take((size_t)P->in_dim * D);
    P->p_c_scale = take(D); P->p_c_shift = take(D); P->p_c_alpha = take(D);
    for (int l = 0; l < d.layers; ++l) {
      auto& L = P->layer[l];
      L.wqkv = take((size_t)D * 3 * D); L.bqkv = take(3 * D);
      L.wo = take((size_t)D * D); L.bo = take(D);
      L.w1 = take((size_t)D * d.ffn); L.b1 = take(d.ffn);
      L.w2 = take((size_t)d.ffn * D); L.b2 = take(D);
      L.ln1w = take(D); L.ln1b = take(D); L.ln2w = take(D); L.ln2b = take(D);
    }
    P->p_ew = take((size_t)D * d.num_event_classes); P->p_eb = take(d.num_event_classes);
    int cin = D;
    for (int i = 0; i < d.num_decoder_layers; ++i) {
      auto& X = P->dec[i];
      X.cin = cin; X.cout = d.decoder_widths[i];
      if (X.cout < 1 || X.cout > D) return false;
      X.w = take((size_t)cin * X.cout); X.scale = take(X.cout); X.shift = take(X.cout); X.alpha = take(X.cout);
      cin = X.cout;
    }
    P->p_ow = take((size_t)cin * d.num_prong_classes); P->p_ob = take(d.num_prong_classes);
    P->packed_bytes = p;
    return true;
  }
};

// device view of the packed block
struct SeqDev {
  const float *pos, *cw, *c_scale, *c_shift, *c_alpha;
  struct Layer { const float *wqkv, *bqkv, *wo, *bo, *w1, *b1, *w2, *b2, *ln1w, *ln1b, *ln2w, *ln2b; } layer[16];
  const float *ew, *eb;
  struct Dec { const float *w, *scale, *shift, *alpha; int cin, cout; } dec[TCVN_MAX_DECODER_LAYERS];
  const float *ow, *ob;
  int D, heads, layers, ffn, pixel_dim, feature_dim, position_dim, in_dim, n_event_classes, n_prong_classes, n_dec;
  float ln_eps;
};

enum { EPI_BIAS = 0, EPI_AFFINE_PRELU = 1, EPI_BIAS_GELU = 2, EPI_NONE = 3 };

// y[s][n] = epi( sum_k x[s][k] * Wt[k][n] ) for s < S; x, y in shared memory; Wt [K][N] in global (n contiguous).
// A thread owns output column n for 16 rows at a time (a <= 16-token sequence reads every weight once); k advances four
// at a time: four weight loads in flight (eight with the unroll) and one 16-byte broadcast read of x per row - the loop is
// a chain of L2 latencies otherwise.  The sum over k keeps its ascending order.
template <int EPI>
__device__ void linear_rows(const float* __restrict__ Wt, const float* __restrict__ p0, const float* __restrict__ p1,
                            const float* __restrict__ p2, const float* x, int ldx, int K, int N, int S, float* y,
                            int ldy) {
  constexpr int R = 16;
  const bool vec = ((ldx | K) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    for (int s0 = 0; s0 < S; s0 += R) {
      float acc[R];
#pragma unroll
      for (int i = 0; i < R; ++i) acc[i] = 0.f;
      if (vec) {
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
          const float w0 = __ldg(Wt + (size_t)k * N + n), w1 = __ldg(Wt + (size_t)(k + 1) * N + n);
          const float w2 = __ldg(Wt + (size_t)(k + 2) * N + n), w3 = __ldg(Wt + (size_t)(k + 3) * N + n);
#pragma unroll
          for (int i = 0; i < R; ++i) {   // rows >= S are scratch
            const float4 xv = *reinterpret_cast<const float4*>(x + (s0 + i) * ldx + k);
            acc[i] = fmaf(xv.x, w0, acc[i]);
            acc[i] = fmaf(xv.y, w1, acc[i]);
            acc[i] = fmaf(xv.z, w2, acc[i]);
            acc[i] = fmaf(xv.w, w3, acc[i]);
          }
        }
      } else {
        for (int k = 0; k < K; ++k) {
          const float w = __ldg(Wt + (size_t)k * N + n);
#pragma unroll
          for (int i = 0; i < R; ++i) acc[i] = fmaf(x[(s0 + i) * ldx + k], w, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        if (s0 + i >= S) break;
        float v = acc[i];
        if (EPI == EPI_BIAS) v += __ldg(p0 + n);
        else if (EPI == EPI_AFFINE_PRELU) v = prelu(fmaf(v, __ldg(p0 + n), __ldg(p1 + n)), __ldg(p2 + n));
        else if (EPI == EPI_BIAS_GELU) {
          v += __ldg(p0 + n);
          v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        }
        y[(s0 + i) * ldy + n] = v;
      }
    }
  }
}

// x[s][:] = LayerNorm(x[s][:] + r[s][:]) ; one warp per token, D <= 128
__device__ void add_layernorm(float* x, const float* r, int S, int D, const float* __restrict__ w,
                              const float* __restrict__ b, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int s = warp; s < S; s += nwarps) {
    float v[4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      v[i] = j < D ? x[s * D + j] + r[s * D + j] : 0.f;
      sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)D;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      const float dlt = j < D ? v[i] - mean : 0.f;
      var += dlt * dlt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)D + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      if (j < D) x[s * D + j] = (v[i] - mean) * rstd * __ldg(w + j) + __ldg(b + j);
    }
  }
}

__global__ void __launch_bounds__(kSeqThreads, 2) seq_forward_kernel(const SeqDev P, int stages,
                                                                  const float* __restrict__ event_emb,
                                                                  const float* __restrict__ prong_emb,
                                                                  const uint8_t* __restrict__ event_mask,
                                                                  const uint8_t* __restrict__ prong_mask,
                                                                  const int* __restrict__ prong_offset, int B, int L,
                                                                  float* tokens, float* hidden, float* event_logits,
                                                                  float* prong_logits) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x;
  const int S = 1 + L;
  const int D = P.D;
  // buffers are sized by the batch's sequence length (rows rounded up to the 16-row register tiles of linear_rows),
  // so several events fit one SM
  const int SR = (S + 15) & ~15;
  float* x = sm;                          // [SR][D]     current hidden state
  float* big = x + SR * D;                // [SR][3D]    token inputs / qkv / ffn hidden
  float* ctx = big + SR * 3 * D;          // [SR][D]     attention context / sub-layer output
  float* xin = ctx + SR * D;              // [SR][in_dim] token-assembly inputs
  __shared__ int valid[kSeqMax];
  __shared__ int prow[kSeqMax];
  if (threadIdx.x < kSeqMax) {
    const int s = threadIdx.x;
    int v = 0, row = -1;
    if (s == 0) v = event_mask ? event_mask[b] != 0 : 1;
    else if (s < S) {
      v = prong_mask[(size_t)b * L + s - 1] != 0;
      int rank = 0;
      for (int t = 0; t < s - 1; ++t) rank += prong_mask[(size_t)b * L + t] != 0;
      row = prong_offset[b] + rank;
    }
    valid[s] = v;
    prow[s] = row;
  }
  __syncthreads();

  if (stages & TCVN_SEQ_TOKENS) {
    // rows: event = [event_emb (pixel+feature) | pos] ; prong = [0 (feature) | prong_emb (pixel) | pos]
    const int ev_dim = P.pixel_dim + P.feature_dim;
    for (int i = threadIdx.x; i < S * P.in_dim; i += blockDim.x) {
      const int s = i / P.in_dim, k = i - s * P.in_dim;
      float v = 0.f;
      if (s == 0) v = k < ev_dim ? event_emb[(size_t)b * ev_dim + k] : __ldg(P.pos + k - ev_dim);
      else if (valid[s]) {
        if (k < P.feature_dim) v = 0.f;
        else if (k < ev_dim) v = prong_emb[(size_t)prow[s] * P.pixel_dim + k - P.feature_dim];
        else v = __ldg(P.pos + k - ev_dim);
      }
      xin[s * P.in_dim + k] = v;
    }
    __syncthreads();
    linear_rows<EPI_AFFINE_PRELU>(P.cw, P.c_scale, P.c_shift, P.c_alpha, xin, P.in_dim, P.in_dim, D, S, x, D);
    __syncthreads();
    // padded slots are zero rows (masked_pad_1d scatters into zeros, packed_data.py:69-76)
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
      const int s = i / D;
      if (s > 0 && !valid[s]) x[i] = 0.f;
      if (tokens) tokens[(size_t)b * S * D + i] = x[i];
    }
    __syncthreads();
  } else if (stages & TCVN_SEQ_ENCODER) {
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) x[i] = tokens[(size_t)b * S * D + i];
    __syncthreads();
  }

  if (stages & TCVN_SEQ_ENCODER) {
    for (int i = threadIdx.x; i < S * D; i += blockDim.x)
      if (!valid[i / D]) x[i] = 0.f;  // hidden * mask, prong_custom_bert_encoder.py:66
    __syncthreads();
    const int dh = D / P.heads;
    const float qscale = rsqrtf((float)dh);
    for (int l = 0; l < P.layers; ++l) {
      const SeqDev::Layer& W = P.layer[l];
      linear_rows<EPI_BIAS>(W.wqkv, W.bqkv, nullptr, nullptr, x, D, D, 3 * D, S, big, 3 * D);
      __syncthreads();
      for (int i = threadIdx.x; i < P.heads * S; i += blockDim.x) {
        const int h = i / S, s = i - h * S;
        const float* q = big + s * 3 * D + h * dh;
        float sc[kSeqMax];
        float mx = -INFINITY;
        for (int t = 0; t < S; ++t) {
          float dot = -INFINITY;
          if (valid[t]) {
            const float* kk = big + t * 3 * D + D + h * dh;
            dot = 0.f;
            for (int j = 0; j < dh; ++j) dot = fmaf(q[j] * qscale, kk[j], dot);
          }
          sc[t] = dot;
          mx = fmaxf(mx, dot);
        }
        float den = 0.f;
        for (int t = 0; t < S; ++t) {
          const float e = valid[t] ? expf(sc[t] - mx) : 0.f;
          sc[t] = e;
          den += e;
        }
        const float inv = 1.f / den;
        for (int j = 0; j < dh; ++j) {
          float o = 0.f;
          for (int t = 0; t < S; ++t) o = fmaf(sc[t], big[t * 3 * D + 2 * D + h * dh + j], o);
          ctx[s * D + h * dh + j] = o * inv;
        }
      }
      __syncthreads();
      float* proj = big;  // qkv is dead now
      linear_rows<EPI_BIAS>(W.wo, W.bo, nullptr, nullptr, ctx, D, D, D, S, proj, D);
      __syncthreads();
      add_layernorm(x, proj, S, D, W.ln1w, W.ln1b, P.ln_eps);
      __syncthreads();
      float* hmid = big;                // [S][ffn]
      linear_rows<EPI_BIAS_GELU>(W.w1, W.b1, nullptr, nullptr, x, D, D, P.ffn, S, hmid, P.ffn);
      __syncthreads();
      linear_rows<EPI_BIAS>(W.w2, W.b2, nullptr, nullptr, hmid, P.ffn, P.ffn, D, S, ctx, D);
      __syncthreads();
      add_layernorm(x, ctx, S, D, W.ln2w, W.ln2b, P.ln_eps);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
      const int s = i / D;
      if (!valid[s]) x[i] = 0.f;  // hidden * mask, prong_custom_bert_encoder.py:73
      if (hidden) hidden[((size_t)s * B + b) * D + (i - s * D)] = x[i];
    }
    __syncthreads();
  } else if (stages & TCVN_SEQ_HEADS) {
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
      const int s = i / D;
      x[i] = hidden[((size_t)s * B + b) * D + (i - s * D)];
    }
    __syncthreads();
  }

  if (stages & TCVN_SEQ_HEADS) {
    linear_rows<EPI_BIAS>(P.ew, P.eb, nullptr, nullptr, x, D, D, P.n_event_classes, 1, ctx, D);
    __syncthreads();
    for (int i = threadIdx.x; i < P.n_event_classes; i += blockDim.x)
      event_logits[(size_t)b * P.n_event_classes + i] = ctx[i];
    __syncthreads();
    // prong MLP over every slot (padded slots carry a zero vector, so they yield the constant row
    // the reference shows for them)
    const float* cur = x + D;
    int ld = D;
    float* bufs[2] = {big, ctx};
    int which = 0;
    for (int i = 0; i < P.n_dec; ++i) {
      const SeqDev::Dec& X = P.dec[i];
      linear_rows<EPI_AFFINE_PRELU>(X.w, X.scale, X.shift, X.alpha, cur, ld, X.cin, X.cout, L, bufs[which], D);
      __syncthreads();
      cur = bufs[which];
      ld = D;
      which ^= 1;
    }
    const int cin = P.n_dec ? P.dec[P.n_dec - 1].cout : D;
    linear_rows<EPI_BIAS>(P.ow, P.ob, nullptr, nullptr, cur, ld, cin, P.n_prong_classes, L, bufs[which], D);
    __syncthreads();
    for (int i = threadIdx.x; i < L * P.n_prong_classes; i += blockDim.x) {
      const int s = i / P.n_prong_classes, c = i - s * P.n_prong_classes;
      prong_logits[((size_t)b * L + s) * P.n_prong_classes + c] = bufs[which][s * D + c];
    }
  }
}

// exclusive prefix sum of valid prongs per event (single CTA; B is a few hundred)
__global__ void prong_offsets_kernel(const uint8_t* __restrict__ prong_mask, int B, int L, int* __restrict__ offsets) {
  __shared__ int partial[1024];
  const int tid = threadIdx.x;
  const int per = (B + blockDim.x - 1) / blockDim.x;
  const int b0 = tid * per, b1 = min(B, b0 + per);
  int s = 0;
  for (int b = b0; b < b1; ++b)
    for (int l = 0; l < L; ++l) s += prong_mask[(size_t)b * L + l] != 0;
  partial[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < blockDim.x; ++i) { const int v = partial[i]; partial[i] = run; run += v; }
    offsets[B] = run;
  }
  __syncthreads();
  int run = partial[tid];
  for (int b = b0; b < b1; ++b) {
    offsets[b] = run;
    for (int l = 0; l < L; ++l) run += prong_mask[(size_t)b * L + l] != 0;
  }
}

static SeqDev make_dev(const SeqPlan& P, const char* pk) {
  SeqDev D{};
  auto f = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  D.pos = f(P.p_pos); D.cw = f(P.p_cw); D.c_scale = f(P.p_c_scale); D.c_shift = f(P.p_c_shift); D.c_alpha = f(P.p_c_alpha);
  for (int l = 0; l < P.d.layers; ++l) {
    const auto& s = P.layer[l];
    D.layer[l] = {f(s.wqkv), f(s.bqkv), f(s.wo), f(s.bo), f(s.w1), f(s.b1), f(s.w2), f(s.b2),
                  f(s.ln1w), f(s.ln1b), f(s.ln2w), f(s.ln2b)};
  }
  D.ew = f(P.p_ew); D.eb = f(P.p_eb);
  for (int i = 0; i < P.d.num_decoder_layers; ++i)
    D.dec[i] = {f(P.dec[i].w), f(P.dec[i].scale), f(P.dec[i].shift), f(P.dec[i].alpha), P.dec[i].cin, P.dec[i].cout};
  D.ow = f(P.p_ow); D.ob = f(P.p_ob);
  D.D = P.d.hidden; D.heads = P.d.heads; D.layers = P.d.layers; D.ffn = P.d.ffn; D.pixel_dim = P.d.pixel_dim;
  D.feature_dim = P.d.feature_dim; D.position_dim = P.d.position_dim; D.in_dim = P.in_dim;
  D.n_event_classes = P.d.num_event_classes; D.n_prong_classes = P.d.num_prong_classes; D.n_dec = P.d.num_decoder_layers;
  D.ln_eps = P.d.ln_eps;
  return D;
}

}  // namespace tcvn

using namespace tcvn;

extern "C" size_t tcvn_seq_packed_bytes(const tcvn_seq_desc* d) {
  SeqPlan P;
  if (!d || !SeqPlan::build(*d, &P)) { set_error("seq: bad descriptor"); return 0; }
  return P.packed_bytes;
}

extern "C" size_t tcvn_seq_workspace_bytes(const tcvn_seq_desc*, int n_events, int) {
  return ((size_t)(n_events > 0 ? n_events : 0) + 1) * sizeof(int) + 256;
}

extern "C" int tcvn_seq_pack(const tcvn_seq_desc* d, const float* position, const float* combined, const float* encoder,
                             const float* event_decoder, const float* prong_decoder, void* packed, size_t packed_bytes,
                             tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && position && combined && encoder && event_decoder && prong_decoder && packed, "seq_pack: null pointer");
  SeqPlan P;
  TCVN_CHECK_ARG(SeqPlan::build(*d, &P), "seq_pack: bad descriptor");
  if (packed_bytes < P.packed_bytes)
    return fail(TCVN_ERR_WORKSPACE, "seq_pack: packed buffer %zu < %zu bytes", packed_bytes, P.packed_bytes);
  char* pk = static_cast<char*>(packed);
  cudaStream_t st = stream;
  const int NOGAP = 1 << 30;
  const int D = d->hidden;
  auto fp = [&](size_t off) { return reinterpret_cast<float*>(pk + off); };
  // transposes: reference Linear weights are [out][in]; the kernel reads Wt[in][out]
  auto tr = [&](const float* src, int n_out, int k_in, size_t off) {
    return repack(src, n_out, k_in, 1, NOGAP, NOGAP, k_in, n_out, false, false, pk + off, st);
  };
  auto bnat = [&](int64_t base, int c) { BnArena r; r.c = c; r.w = base; r.b = base + c; r.rm = base + 2 * c; r.rv = base + 3 * c; r.alpha = base + 4 * c; return r; };
  TCVN_TRY(pad_copy(position, d->position_dim, fp(P.p_pos), d->position_dim, st));
  // combined_embedding: linear.weight [D][in], norm (w,b,rm,rv), activation.weight
  TCVN_TRY(tr(combined, D, P.in_dim, P.p_cw));
  TCVN_TRY(fold(combined, bnat((int64_t)D * P.in_dim, D), nullptr, NOGAP, NOGAP, D, d->bn_eps, pk, P.p_c_scale,
                P.p_c_shift, P.p_c_alpha, st));
  int64_t a = 0;
  for (int l = 0; l < d->layers; ++l) {
    const auto& L = P.layer[l];
    TCVN_TRY(tr(encoder + a, 3 * D, D, L.wqkv)); a += (int64_t)3 * D * D;
    TCVN_TRY(pad_copy(encoder + a, 3 * D, fp(L.bqkv), 3 * D, st)); a += 3 * D;
    TCVN_TRY(tr(encoder + a, D, D, L.wo)); a += (int64_t)D * D;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.bo), D, st)); a += D;
    TCVN_TRY(tr(encoder + a, d->ffn, D, L.w1)); a += (int64_t)d->ffn * D;
    TCVN_TRY(pad_copy(encoder + a, d->ffn, fp(L.b1), d->ffn, st)); a += d->ffn;
    TCVN_TRY(tr(encoder + a, D, d->ffn, L.w2)); a += (int64_t)D * d->ffn;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.b2), D, st)); a += D;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.ln1w), D, st)); a += D;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.ln1b), D, st)); a += D;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.ln2w), D, st)); a += D;
    TCVN_TRY(pad_copy(encoder + a, D, fp(L.ln2b), D, st)); a += D;
  }
  TCVN_TRY(tr(event_decoder, d->num_event_classes, D, P.p_ew));
  TCVN_TRY(pad_copy(event_decoder + (int64_t)d->num_event_classes * D, d->num_event_classes, fp(P.p_eb),
                    d->num_event_classes, st));
  a = 0;
  for (int i = 0; i < d->num_decoder_layers; ++i) {
    const auto& X = P.dec[i];
    TCVN_TRY(tr(prong_decoder + a, X.cout, X.cin, X.w));
    const float* bias = prong_decoder + a + (int64_t)X.cout * X.cin;
    TCVN_TRY(fold(prong_decoder, bnat(a + (int64_t)X.cout * X.cin + X.cout, X.cout), bias, NOGAP, NOGAP, X.cout,
                  d->bn_eps, pk, X.scale, X.shift, X.alpha, st));
    a += (int64_t)X.cout * X.cin + X.cout + 5 * X.cout;
  }
  const int cin = d->num_decoder_layers ? P.dec[d->num_decoder_layers - 1].cout : D;
  TCVN_TRY(tr(prong_decoder + a, d->num_prong_classes, cin, P.p_ow));
  TCVN_TRY(pad_copy(prong_decoder + a + (int64_t)d->num_prong_classes * cin, d->num_prong_classes, fp(P.p_ob),
                    d->num_prong_classes, st));
  return TCVN_OK;
}

extern "C" int tcvn_seq_forward(const tcvn_seq_desc* d, const void* packed, int stages, const float* event_embedding,
                                const float* prong_embedding, const uint8_t* event_mask, const uint8_t* prong_mask,
                                int n_events, int max_prongs, float* tokens, float* hidden, float* event_logits,
                                float* prong_logits, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && packed && workspace && (prong_mask || max_prongs == 0), "seq_forward: null pointer");
  TCVN_CHECK_ARG(stages > 0 && stages < 8, "seq_forward: bad stage mask %d", stages);
  TCVN_CHECK_ARG(stages != (TCVN_SEQ_TOKENS | TCVN_SEQ_HEADS), "seq_forward: stages must be contiguous");
  TCVN_CHECK_ARG(n_events >= 0 && max_prongs >= 0, "seq_forward: negative size");
  if (n_events == 0) return TCVN_OK;
  if (1 + max_prongs > kSeqMax)
    return fail(TCVN_ERR_UNSUPPORTED, "seq_forward: %d prong slots (kernel holds 1+%d tokens)", max_prongs, kSeqMax - 1);
  SeqPlan P;
  TCVN_CHECK_ARG(SeqPlan::build(*d, &P), "seq_forward: bad descriptor");
  if (workspace_bytes < tcvn_seq_workspace_bytes(d, n_events, max_prongs))
    return fail(TCVN_ERR_WORKSPACE, "seq_forward: workspace too small");
  const bool tok = stages & TCVN_SEQ_TOKENS, enc = stages & TCVN_SEQ_ENCODER, hd = stages & TCVN_SEQ_HEADS;
  if (tok) TCVN_CHECK_ARG(event_embedding && (prong_embedding || max_prongs == 0), "seq_forward: embeddings missing");
  if (tok && !enc) TCVN_CHECK_ARG(tokens, "seq_forward: tokens output missing");
  if (enc && !tok) TCVN_CHECK_ARG(tokens, "seq_forward: tokens input missing");
  if (enc && !hd) TCVN_CHECK_ARG(hidden, "seq_forward: hidden output missing");
  if (hd && !enc) TCVN_CHECK_ARG(hidden, "seq_forward: hidden input missing");
  if (hd) TCVN_CHECK_ARG(event_logits && (prong_logits || max_prongs == 0), "seq_forward: logits output missing");
  int* offsets = static_cast<int*>(workspace);
  cudaStream_t st = stream;
  prong_offsets_kernel<<<1, 256, 0, st>>>(prong_mask, n_events, max_prongs, offsets);
  TCVN_LAUNCH_CHECK();
  const int SR = (1 + max_prongs + 15) & ~15;
  const size_t smem = ((size_t)SR * d->hidden * 5 + (size_t)SR * P.in_dim) * sizeof(float);
  TCVN_CUDA(cudaFuncSetAttribute(seq_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // latency-bound, one CTA per event: ask for the full shared-memory carve-out so that two CTAs are resident per SM
  TCVN_CUDA(cudaFuncSetAttribute(seq_forward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  seq_forward_kernel<<<n_events, kSeqThreads, smem, st>>>(make_dev(P, static_cast<const char*>(packed)), stages,
                                                          event_embedding, prong_embedding, event_mask, prong_mask,
                                                          offsets, n_events, max_prongs, tokens, hidden, event_logits,
                                                          prong_logits);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
