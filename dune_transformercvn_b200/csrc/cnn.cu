// DenseNet pixel-map embedding, eval-mode forward: the layer walk.
// Follows transformercvn/network/layers/dense_net.py:111-167 (DenseNet.features / condense /
// output_block) with the dataflow re-designed for in-place concat buffers (no torch.cat, dense_net.py:45):
//   stem -> pool -> blk[0][:, 0:64]
//   layer i of block b:  mid = PReLU2(BN2(conv1(PReLU1(BN1(blk[b][:, :k_i])))))      (one fused GEMM)
//                        blk[b][:, k_i : k_i+32] = conv2_3x3(mid)                     (one shifted GEMM)
//   transition b:        blk[b+1][:, 0:c/2] = conv1x1(avgpool2(PReLU(BN(blk[b]))))    (pool first: 4x fewer FLOPs)
//   tail:                embedding = PReLU(BN1d(Linear(mean_hw(PReLU(BN(blk[last]))))))
// Images are processed in per-block chunks (plan.h) so a block's concat buffer stays L2-resident between its
// layers while the small late blocks still get enough rows per launch to fill the machine.
#include <stdlib.h>

#include "kernels.h"
#include "plan.h"
#include "umma.h"

using namespace tcvn;



namespace {

inline const float* pf(const char* packed, size_t off) { return reinterpret_cast<const float*>(packed + off); }

struct Walk {
  const CnnPlan& P;
  const char* pk;
  const float* pixels;  // whole batch
  char* ws;
  cudaStream_t st;
  bool f32;
  // COO-direct ingest (pixels == nullptr): the stem reads the hit list itself
  const int32_t* coords = nullptr;
  const void* values = nullptr;
  bool values_u8 = false;
  float divisor = 0.f;
  void* bins = nullptr;  // scratch of the binned COO stem (present when the caller sized the workspace with the nnz-aware query)
  int n_binned = 0;      // images / hits the bins were built for (the whole call)
  long long nnz_binned = 0;

  int stem(int i0, int n) {
    const tcvn_cnn_desc& d = P.d;
    const size_t img_floats = (size_t)d.in_channels * d.height * d.width;
    const BlockPlan& B0 = P.blocks[0];
    if (pixels == nullptr && bins != nullptr)
      return launch_stem_coo_binned(i0, n, n_binned, nnz_binned, d.in_channels, d.height, d.width, pf(pk, P.p_w0),
                                    pf(pk, P.p_s_scale), pf(pk, P.p_s_shift), pf(pk, P.p_s_alpha), d.init_features,
                                    ws + B0.ws_blk, B0.ld, B0.H, B0.W, f32, bins, st);
    if (pixels == nullptr)
      return launch_stem_coo(coords, values, values_u8, reinterpret_cast<const long long*>(ws + P.ws_hitofs), i0, divisor,
                             n, d.in_channels, d.height, d.width, pf(pk, P.p_w0), pf(pk, P.p_s_scale),
                             pf(pk, P.p_s_shift), pf(pk, P.p_s_alpha), d.init_features, ws + B0.ws_blk, B0.ld, B0.H,
                             B0.W, f32, st);
    return launch_stem(pixels + (size_t)i0 * img_floats, n, d.in_channels, d.height, d.width, pf(pk, P.p_w0),
                       pf(pk, P.p_s_scale), pf(pk, P.p_s_shift), pf(pk, P.p_s_alpha), d.init_features, ws + B0.ws_blk,
                       B0.ld, B0.H, B0.W, f32, st);
  }

  int dense_block(const BlockPlan& B, int n) {
    const tcvn_cnn_desc& d = P.d;
    void* blk = ws + B.ws_blk;
    void* mid = ws + P.ws_mid;
    const long long rows = (long long)n * B.R;
    for (const LayerPlan& L : B.layers) {
      if (f32) {
        GemmArgs g{};
        g.A = blk; g.lda = B.ld; g.m_total = rows; g.K = L.kphys; g.taps = 1; g.tap_off[0] = 0;
        g.W = pf(pk, L.p_w1); g.N = P.mid;
        g.a_scale = pf(pk, L.p_a_scale); g.a_shift = pf(pk, L.p_a_shift); g.a_alpha = pf(pk, L.p_a_alpha);
        g.o_scale = pf(pk, L.p_o_scale); g.o_shift = pf(pk, L.p_o_shift); g.o_alpha = pf(pk, L.p_o_alpha);
        g.out = mid; g.ldo = P.mid; g.out_col0 = 0; g.ring_Hp = B.Hp; g.ring_Wp = B.Wp;
        g.a_is_f32 = true; g.out_is_f32 = true;
        TCVN_TRY(launch_simt_gemm(g, st));
        GemmArgs c{};
        c.A = mid; c.lda = P.mid; c.m_total = rows; c.K = P.mid; c.taps = 9;
        for (int t = 0; t < 9; ++t) c.tap_off[t] = (t / 3 - 1) * B.Wp + (t % 3 - 1);
        c.W = pf(pk, L.p_w2); c.N = d.growth;
        c.o_shift = pf(pk, L.p_b2);
        c.out = blk; c.ldo = B.ld; c.out_col0 = L.kphys; c.ring_Hp = B.Hp; c.ring_Wp = B.Wp;
        c.a_is_f32 = true; c.out_is_f32 = true;
        TCVN_TRY(launch_simt_gemm(c, st));
      } else {
        TCVN_TRY(umma_dense_layer(P, B, L, pk, blk, mid, rows, st));
      }
    }
    return TCVN_OK;
  }

  // transition after block b over n images -> images [j, j+n) of block b+1's buffer
  int transition(int b, int n, int j) {
    const BlockPlan& B = P.blocks[b];
    const BlockPlan& Nx = P.blocks[b + 1];
    void* pool = ws + P.ws_pool;
    TCVN_TRY(launch_act_pool2(ws + B.ws_blk, n, B.H, B.W, B.ld, B.ctot, pf(pk, B.p_t_scale), pf(pk, B.p_t_shift),
                              pf(pk, B.p_t_alpha), pool, Nx.H, Nx.W, f32, st));
    if (!f32)
      return umma_transition(P, B, Nx, pk, pool, ws + Nx.ws_blk + (size_t)j * Nx.R * Nx.ld * P.esize,
                             (long long)n * Nx.R, st);
    GemmArgs g{};
    g.A = pool; g.lda = B.ctot; g.m_total = (long long)n * Nx.R; g.K = B.ctot; g.taps = 1; g.tap_off[0] = 0;
    g.W = pf(pk, B.p_tw); g.N = B.toutp;
    g.o_shift = pf(pk, B.p_tb);
    g.out = ws + Nx.ws_blk + (size_t)j * Nx.R * Nx.ld * P.esize; g.ldo = Nx.ld; g.out_col0 = 0;
    g.ring_Hp = Nx.Hp; g.ring_Wp = Nx.Wp;
    g.a_is_f32 = f32; g.out_is_f32 = f32;
    return launch_simt_gemm(g, st);
  }

  // fills block b's buffer for images [i0, i0+n) of the batch and runs its dense layers
  int process(int b, int i0, int n) {
    if (b == 0) {
      TCVN_TRY(stem(i0, n));
    } else {
      const int sub = P.blocks[b - 1].chunk;
      for (int j = 0; j < n; j += sub) {
        const int m = n - j < sub ? n - j : sub;
        TCVN_TRY(process(b - 1, i0 + j, m));
        TCVN_TRY(transition(b - 1, m, j));
      }
    }
    return dense_block(P.blocks[b], n);
  }

  int tail(int n, float* embedding) {
    const tcvn_cnn_desc& d = P.d;
    const BlockPlan& last = P.blocks.back();
    float* gap = reinterpret_cast<float*>(ws + P.ws_gap);
    TCVN_TRY(launch_act_gap(ws + last.ws_blk, n, last.H, last.W, last.ld, last.ctot, pf(pk, P.p_f_scale),
                            pf(pk, P.p_f_shift), pf(pk, P.p_f_alpha), gap, f32, st));
    GemmArgs g{};
    g.A = gap; g.lda = last.ctot; g.m_total = n; g.K = last.ctot; g.taps = 1; g.tap_off[0] = 0;
    g.W = pf(pk, P.p_lw); g.N = d.out_features;
    g.o_scale = pf(pk, P.p_lo_scale); g.o_shift = pf(pk, P.p_lo_shift); g.o_alpha = pf(pk, P.p_lo_alpha);
    g.out = embedding; g.ldo = d.out_features; g.out_col0 = 0;
    g.a_is_f32 = true; g.out_is_f32 = true;
    return launch_simt_gemm(g, st);
  }
};

}  // namespace

extern "C" int tcvn_cnn_forward(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed, const float* pixels,
                                int n_images, float* embedding, void* workspace, size_t workspace_bytes,
                                tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && packed && workspace, "cnn_forward: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_FP32 || prec == TCVN_BF16, "cnn_forward: unknown precision");
  TCVN_CHECK_ARG(n_images >= 0, "cnn_forward: negative image count");
  if (n_images == 0) return TCVN_OK;
  TCVN_CHECK_ARG(pixels && embedding, "cnn_forward: null pointer");
  CnnPlan P;
  TCVN_CHECK_ARG(CnnPlan::build(*d, prec, n_images, &P, true), "cnn_forward: bad descriptor");
  if (workspace_bytes < P.ws_bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_forward: workspace %zu < %zu bytes", workspace_bytes, P.ws_bytes);
  Walk w{P, static_cast<const char*>(packed), pixels, static_cast<char*>(workspace), stream, prec == TCVN_FP32};
  const int last = (int)P.blocks.size() - 1;
  const int top = P.blocks[last].chunk;
  for (int i0 = 0; i0 < n_images; i0 += top) {
    const int n = n_images - i0 < top ? n_images - i0 : top;
    TCVN_TRY(w.process(last, i0, n));
    TCVN_TRY(w.tail(n, embedding + (size_t)i0 * d->out_features));
  }
  return TCVN_OK;
}

extern "C" int tcvn_cnn_forward_sparse(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed,
                                       const int32_t* coords, const void* values, tcvn_value_dtype value_dtype,
                                       int64_t nnz, float divisor, int n_images, float* embedding, void* workspace,
                                       size_t workspace_bytes, tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && packed && workspace, "cnn_forward_sparse: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_FP32 || prec == TCVN_BF16, "cnn_forward_sparse: unknown precision");
  TCVN_CHECK_ARG(value_dtype == TCVN_VAL_F32 || value_dtype == TCVN_VAL_U8, "cnn_forward_sparse: unknown value dtype");
  TCVN_CHECK_ARG(n_images >= 0 && nnz >= 0, "cnn_forward_sparse: negative size");
  if (n_images == 0) return TCVN_OK;
  TCVN_CHECK_ARG(embedding && (nnz == 0 || (coords && values)), "cnn_forward_sparse: null pointer");
  CnnPlan P;
  TCVN_CHECK_ARG(CnnPlan::build(*d, prec, n_images, &P, true), "cnn_forward_sparse: bad descriptor");
  if (workspace_bytes < P.ws_bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_forward_sparse: workspace %zu < %zu bytes", workspace_bytes, P.ws_bytes);
  Walk w{P, static_cast<const char*>(packed), nullptr, static_cast<char*>(workspace), stream, prec == TCVN_FP32};
  w.coords = coords; w.values = values; w.values_u8 = value_dtype == TCVN_VAL_U8; w.divisor = divisor;
  {
    const size_t base = align_up(P.ws_bytes, 1024);
    const size_t extra = stem_bins_bytes(n_images, P.blocks[0].H, P.blocks[0].W, nnz);
    if (workspace_bytes >= base + extra && nnz < (1ll << 28) && !getenv("TCVN_STEM_UNBINNED")) w.bins = w.ws + base;
  }
  TCVN_TRY(launch_hit_offsets(coords, nnz, n_images, reinterpret_cast<long long*>(w.ws + P.ws_hitofs), stream));
  if (w.bins) {
    w.n_binned = n_images; w.nnz_binned = nnz;
    TCVN_TRY(launch_stem_bin(coords, values, w.values_u8, reinterpret_cast<const long long*>(w.ws + P.ws_hitofs), n_images, nnz,
                             d->in_channels, divisor, d->height, d->width, P.blocks[0].H, P.blocks[0].W, w.bins, stream));
  }
  const int last = (int)P.blocks.size() - 1;
  const int top = P.blocks[last].chunk;
  for (int i0 = 0; i0 < n_images; i0 += top) {
    const int n = n_images - i0 < top ? n_images - i0 : top;
    TCVN_TRY(w.process(last, i0, n));
    TCVN_TRY(w.tail(n, embedding + (size_t)i0 * d->out_features));
  }
  return TCVN_OK;
}

extern "C" size_t tcvn_cnn_workspace_bytes_sparse(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images, int64_t nnz) {
  CnnPlan P;
  if (!d || nnz < 0 || !CnnPlan::build(*d, prec, n_images, &P, true)) { set_error("cnn: bad descriptor"); return 0; }
  return align_up(P.ws_bytes, 1024) + stem_bins_bytes(n_images, P.blocks[0].H, P.blocks[0].W, nnz);
}

extern "C" int tcvn_cnn_run_layer(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed, void* workspace,
                                  size_t workspace_bytes, int n_images, int block, int layer, int which,
                                  tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && packed && workspace, "cnn_run_layer: null pointer");
  TCVN_CHECK_ARG(prec == TCVN_BF16, "cnn_run_layer: only the tcgen05 path has separately launchable layer kernels");
  CnnPlan P;
  TCVN_CHECK_ARG(CnnPlan::build(*d, prec, n_images, &P, true), "cnn_run_layer: bad descriptor");
  if (workspace_bytes < P.ws_bytes)
    return fail(TCVN_ERR_WORKSPACE, "cnn_run_layer: workspace %zu < %zu bytes", workspace_bytes, P.ws_bytes);
  TCVN_CHECK_ARG(block >= 0 && block < (int)P.blocks.size(), "cnn_run_layer: no block %d", block);
  const BlockPlan& B = P.blocks[block];
  TCVN_CHECK_ARG(layer >= 0 && layer < (int)B.layers.size(), "cnn_run_layer: no layer %d", layer);
  TCVN_CHECK_ARG(n_images >= 1 && n_images <= B.chunk, "cnn_run_layer: %d images (block chunk is %d)", n_images, B.chunk);
  char* ws = static_cast<char*>(workspace);
  return umma_dense_layer_part(P, B, B.layers[layer], static_cast<const char*>(packed), ws + B.ws_blk, ws + P.ws_mid,
                               (long long)n_images * B.R, which, stream);
}

extern "C" int tcvn_cnn_read_stage(const tcvn_cnn_desc* d, tcvn_precision prec, const void* workspace, int n_images,
                                   int stage, float* out, int32_t* channels, int32_t* h, int32_t* w,
                                   tcvn_stream_t stream) {
  TCVN_CHECK_ARG(d && workspace && channels && h && w, "cnn_read_stage: null pointer");
  CnnPlan P;
  TCVN_CHECK_ARG(CnnPlan::build(*d, prec, n_images, &P, true), "cnn_read_stage: bad descriptor");
  TCVN_CHECK_ARG(n_images <= P.chunk, "cnn_read_stage: only the last chunk (%d images) is still in the workspace", P.chunk);
  TCVN_CHECK_ARG(stage >= 0 && stage <= 2 * (int)P.blocks.size() - 1, "cnn_read_stage: no stage %d", stage);
  const int b = stage == 0 ? 0 : (stage % 2 == 1 ? (stage - 1) / 2 : stage / 2);
  const BlockPlan& B = P.blocks[b];
  // stage 0 and transition outputs are the first c0 channels of a block buffer; block outputs are all of it
  const int c = (stage % 2 == 1) ? B.clog : B.c0;
  *channels = c; *h = B.H; *w = B.W;
  if (out == nullptr) return TCVN_OK;
  return launch_ring_to_nchw(static_cast<const char*>(workspace) + B.ws_blk, n_images, B.H, B.W, B.ld, c, B.c0,
                             B.c0p - B.c0, out, prec == TCVN_FP32, stream);
}
