// Loss and validation metrics of the training / validation step on the device, one launch each, no host sync.
//
// Reference arithmetic (transformercvn/network/trainers/neutrino_full_base_trainer.py):
//   loss            :148-160  focal loss  -log p_t * (1 - p_t)^gamma, mean over rows (gamma == 0: cross entropy)
//   training_step   :162-192  prong rows selected by target >= 0 (masked_select = one host sync in the reference),
//                             total = event_scale * event_loss + prong_scale * prong_loss, argmax accuracies
//   validation_step :194-209  softmax probabilities of the event rows and of the selected prong rows, accuracy state
// The kernel is one CTA: the batch has at most a few thousand rows of 4 / 8 classes, and a fixed thread -> row
// assignment with a fixed reduction tree makes the loss bit-reproducible run to run.
#include "common.cuh"

namespace tcvn {

constexpr int kLossThreads = 512;
constexpr int kMaxClasses = 32;

struct LossArgs {
  const float* ev_logits; const long long* ev_targets; int B, E;
  const float* pr_logits; const long long* pr_targets; int L, P;
  long long pr_sb, pr_sl;   // element strides of prong_logits[b][l][:] (the network returns a transposed view)
  float gamma, ev_scale, pr_scale;
  float* out;               // [8]
  float* d_ev; float* d_pr; // contiguous (B, E) / (B, L, P)
};

// block-wide sum of three doubles in a fixed order; result valid in every thread
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* sh /* [3][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __syncthreads();
  if (lane == 0) { sh[warp] = a; sh[32 + warp] = b; sh[64 + warp] = c; }
  __syncthreads();
  a = 0.0; b = 0.0; c = 0.0;
  for (int w = 0; w < nw; ++w) { a += sh[w]; b += sh[32 + w]; c += sh[64 + w]; }
}

// one row: loss, un-normalised logit gradient, argmax hit
__device__ __forceinline__ void focal_row(const float* z, int C, int t_in, float gamma, float* g, float& loss, bool& hit) {
  const int t = t_in < 0 ? 0 : (t_in >= C ? C - 1 : t_in);   // a target outside [0, C) poisons the loss below (one_hot raises)
  float mx = z[0]; int am = 0;
  for (int j = 1; j < C; ++j) if (z[j] > mx) { mx = z[j]; am = j; }   // first maximum, like torch.argmax
  float e[kMaxClasses], s = 0.f;
  for (int j = 0; j < C; ++j) { e[j] = expf(z[j] - mx); s += e[j]; }
  const float inv = 1.f / s;
  const float logp = z[t] - mx - logf(s);
  const float pt = e[t] * inv;
  const float q = 1.f - pt;
  float w, coef;   // w = q^gamma ;  dloss/dz_j = coef * (delta_tj - p_j)
  if (gamma == 0.f) { w = 1.f; coef = -1.f; }
  else if (gamma == 1.f) { w = q; coef = pt * logp - q; }
  else { w = powf(q, gamma); coef = gamma * powf(q, gamma - 1.f) * pt * logp - w; }
  loss = t == t_in ? -logp * w : __int_as_float(0x7fc00000);
  hit = am == t_in;
  if (g != nullptr)
    for (int j = 0; j < C; ++j) g[j] = coef * ((j == t ? 1.f : 0.f) - e[j] * inv);
}

__global__ void __launch_bounds__(kLossThreads) focal_loss_kernel(const LossArgs a) {
  __shared__ double sh[96];
  __shared__ float s_fac[2];
  const int rows_e = a.B, rows_p = a.B * a.L;
  double le = 0.0, ce = 0.0, dummy = 0.0, lp = 0.0, cp = 0.0, np = 0.0;
  for (int r = threadIdx.x; r < rows_e + rows_p; r += kLossThreads) {
    float z[kMaxClasses], loss; bool hit;
    if (r < rows_e) {
      for (int j = 0; j < a.E; ++j) z[j] = a.ev_logits[(long long)r * a.E + j];
      focal_row(z, a.E, (int)a.ev_targets[r], a.gamma, a.d_ev ? a.d_ev + (long long)r * a.E : nullptr, loss, hit);
      le += loss; ce += hit ? 1.0 : 0.0;
    } else {
      const int q = r - rows_e, b = q / a.L, l = q - b * a.L;
      const long long t = a.pr_targets[q];
      float* g = a.d_pr ? a.d_pr + (long long)q * a.P : nullptr;
      if (t < 0) {
        if (g) for (int j = 0; j < a.P; ++j) g[j] = 0.f;
        continue;
      }
      const float* src = a.pr_logits + b * a.pr_sb + l * a.pr_sl;
      for (int j = 0; j < a.P; ++j) z[j] = src[j];
      focal_row(z, a.P, (int)t, a.gamma, g, loss, hit);
      lp += loss; cp += hit ? 1.0 : 0.0; np += 1.0;
    }
  }
  block_sum3(le, ce, dummy, sh);
  block_sum3(lp, cp, np, sh);
  if (threadIdx.x == 0) {
    const float ev_loss = (float)(le / (double)rows_e);
    const float pr_loss = (float)(lp / np);            // no selected prong row: 0/0 = NaN, like mean() of an empty tensor
    a.out[0] = a.ev_scale * ev_loss + a.pr_scale * pr_loss;
    a.out[1] = ev_loss;
    a.out[2] = pr_loss;
    a.out[3] = (float)(ce / (double)rows_e);
    a.out[4] = (float)(cp / np);
    a.out[5] = (float)rows_e;
    a.out[6] = (float)np;
    a.out[7] = 0.f;
    s_fac[0] = a.ev_scale / (float)rows_e;
    s_fac[1] = a.pr_scale / (float)np;
  }
  __syncthreads();
  if (a.d_ev == nullptr) return;
  // every thread rescales the rows it wrote itself
  const float fe = s_fac[0], fp = s_fac[1];
  for (int r = threadIdx.x; r < rows_e + rows_p; r += kLossThreads) {
    if (r < rows_e) {
      for (int j = 0; j < a.E; ++j) a.d_ev[(long long)r * a.E + j] *= fe;
    } else {
      const int q = r - rows_e;
      if (a.pr_targets[q] < 0) continue;
      for (int j = 0; j < a.P; ++j) a.d_pr[(long long)q * a.P + j] *= fp;
    }
  }
}

__global__ void loss_scale_kernel(const float* __restrict__ upstream, const float* __restrict__ g_ev, long long n_ev,
                                  const float* __restrict__ g_pr, long long n_pr, float* __restrict__ o_ev,
                                  float* __restrict__ o_pr) {
  const float u = upstream[0];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ev) o_ev[i] = g_ev[i] * u;
  else if (i < n_ev + n_pr) o_pr[i - n_ev] = g_pr[i - n_ev] * u;
}

struct MetricArgs {
  const float* ev_logits; const long long* ev_targets; int B, E;
  const float* pr_logits; const long long* pr_targets; int L, P;
  long long pr_sb, pr_sl;
  long long* counters;      // [4] += event hits, event rows, prong hits, selected prong rows
  float* ev_prob; float* pr_prob;   // (B, E) / (B, L, P); unselected prong rows are written as zeros
};

__global__ void __launch_bounds__(kLossThreads) metrics_kernel(const MetricArgs a) {
  __shared__ double sh[96];
  const int rows_e = a.B, rows_p = a.B * a.L;
  double ce = 0.0, cp = 0.0, np = 0.0;
  for (int r = threadIdx.x; r < rows_e + rows_p; r += kLossThreads) {
    const bool ev = r < rows_e;
    const int q = ev ? r : r - rows_e;
    const int C = ev ? a.E : a.P;
    float* dst = ev ? a.ev_prob + (long long)q * C : a.pr_prob + (long long)q * C;
    const long long t = ev ? a.ev_targets[q] : a.pr_targets[q];
    if (t < 0) {
      for (int j = 0; j < C; ++j) dst[j] = 0.f;
      continue;
    }
    const float* src = ev ? a.ev_logits + (long long)q * C : a.pr_logits + (q / a.L) * a.pr_sb + (q % a.L) * a.pr_sl;
    float z[kMaxClasses], mx = src[0]; int am = 0;
    z[0] = mx;
    for (int j = 1; j < C; ++j) { z[j] = src[j]; if (z[j] > mx) { mx = z[j]; am = j; } }
    float s = 0.f;
    for (int j = 0; j < C; ++j) { z[j] = expf(z[j] - mx); s += z[j]; }
    const float inv = 1.f / s;
    for (int j = 0; j < C; ++j) dst[j] = z[j] * inv;
    if (ev) ce += am == (int)t ? 1.0 : 0.0;
    else { cp += am == (int)t ? 1.0 : 0.0; np += 1.0; }
  }
  block_sum3(ce, cp, np, sh);
  if (threadIdx.x == 0) {
    a.counters[0] += (long long)ce;
    a.counters[1] += rows_e;
    a.counters[2] += (long long)cp;
    a.counters[3] += (long long)np;
  }
}

}  // namespace tcvn

using namespace tcvn;

extern "C" int tcvn_loss_forward(const float* event_logits, const int64_t* event_targets, int n_events, int event_classes,
                                 const float* prong_logits, const int64_t* prong_targets, int max_prongs, int prong_classes,
                                 int64_t prong_stride_event, int64_t prong_stride_slot, float gamma, float event_scale,
                                 float prong_scale, float* out8, float* d_event_logits, float* d_prong_logits,
                                 tcvn_stream_t stream) {
  TCVN_CHECK_ARG(event_logits && event_targets && prong_logits && prong_targets && out8, "loss_forward: null pointer");
  TCVN_CHECK_ARG(n_events > 0 && max_prongs >= 0 && event_classes >= 1 && event_classes <= kMaxClasses &&
                     prong_classes >= 1 && prong_classes <= kMaxClasses,
                 "loss_forward: n_events %d, max_prongs %d, classes %d / %d (at most %d)", n_events, max_prongs, event_classes,
                 prong_classes, kMaxClasses);
  TCVN_CHECK_ARG((d_event_logits == nullptr) == (d_prong_logits == nullptr), "loss_forward: pass both gradients or neither");
  TCVN_CHECK_ARG(gamma >= 0.f, "loss_forward: gamma %f < 0", gamma);
  LossArgs a;
  a.ev_logits = event_logits; a.ev_targets = reinterpret_cast<const long long*>(event_targets); a.B = n_events; a.E = event_classes;
  a.pr_logits = prong_logits; a.pr_targets = reinterpret_cast<const long long*>(prong_targets); a.L = max_prongs; a.P = prong_classes;
  a.pr_sb = prong_stride_event; a.pr_sl = prong_stride_slot;
  a.gamma = gamma; a.ev_scale = event_scale; a.pr_scale = prong_scale;
  a.out = out8; a.d_ev = d_event_logits; a.d_pr = d_prong_logits;
  focal_loss_kernel<<<1, kLossThreads, 0, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_loss_backward(const float* upstream, const float* g_event, int64_t n_event, const float* g_prong,
                                  int64_t n_prong, float* d_event_logits, float* d_prong_logits, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(upstream && g_event && g_prong && d_event_logits && d_prong_logits && n_event >= 0 && n_prong >= 0,
                 "loss_backward: bad arguments");
  if (n_event + n_prong == 0) return TCVN_OK;
  loss_scale_kernel<<<(unsigned)ceil_div_ll(n_event + n_prong, 256), 256, 0, stream>>>(upstream, g_event, n_event, g_prong,
                                                                                        n_prong, d_event_logits, d_prong_logits);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

extern "C" int tcvn_metrics_update(const float* event_logits, const int64_t* event_targets, int n_events, int event_classes,
                                   const float* prong_logits, const int64_t* prong_targets, int max_prongs, int prong_classes,
                                   int64_t prong_stride_event, int64_t prong_stride_slot, int64_t* counters4,
                                   float* event_prob, float* prong_prob, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(event_logits && event_targets && prong_logits && prong_targets && counters4 && event_prob && prong_prob,
                 "metrics_update: null pointer");
  TCVN_CHECK_ARG(n_events > 0 && max_prongs >= 0 && event_classes >= 1 && event_classes <= kMaxClasses &&
                     prong_classes >= 1 && prong_classes <= kMaxClasses,
                 "metrics_update: n_events %d, max_prongs %d, classes %d / %d (at most %d)", n_events, max_prongs,
                 event_classes, prong_classes, kMaxClasses);
  MetricArgs a;
  a.ev_logits = event_logits; a.ev_targets = reinterpret_cast<const long long*>(event_targets); a.B = n_events; a.E = event_classes;
  a.pr_logits = prong_logits; a.pr_targets = reinterpret_cast<const long long*>(prong_targets); a.L = max_prongs; a.P = prong_classes;
  a.pr_sb = prong_stride_event; a.pr_sl = prong_stride_slot;
  a.counters = reinterpret_cast<long long*>(counters4); a.ev_prob = event_prob; a.pr_prob = prong_prob;
  metrics_kernel<<<1, kLossThreads, 0, stream>>>(a);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
