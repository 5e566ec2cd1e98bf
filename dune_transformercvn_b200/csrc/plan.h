// Host-side plan of one DenseNet pixel-map CNN: geometry of the padded channels-last feature
// maps, offsets into the reference-ordered fp32 arena, offsets into the packed parameter block
// and into the caller's workspace.  Pure arithmetic, no CUDA.
//
// Reference topology: transformercvn/network/layers/dense_net.py:97-162 (walk mirrored from
// dune_transformercvn_b200/params.py::densenet_specs, which is pinned against the reference's
// state_dict by tests/test_params.py).
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../include/tcvn.h"

namespace tcvn {

constexpr int kChanAlign = 8;      // concat-buffer channel slices start on 16-byte (bf16) boundaries
constexpr int kKChunk = 64;        // bf16 K tile = one 128-byte swizzle row

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct BnArena {  // float offsets of one BatchNorm + PReLU pair in the arena
  int64_t w, b, rm, rv, alpha;
  int c;
};

struct LayerPlan {
  int cin;       // logical input channels (C0 + growth*i)
  int kphys;     // physical input channels in the block buffer (C0p + growth*i)
  int kpad;      // kphys rounded up to kKChunk (bf16 weights are zero beyond kphys)
  BnArena norm1, norm2;
  int64_t conv1_w, conv1_b, conv2_w, conv2_b;  // arena offsets
  // packed offsets (bytes)
  size_t p_a_scale, p_a_shift, p_a_alpha;  // [kpad] fp32: BN1+PReLU1 fold over physical channels
  size_t p_w1;                             // fp32: [kphys][mid] ; bf16: [mid][kpad]
  size_t p_o_scale, p_o_shift, p_o_alpha;  // [mid] fp32: conv1 bias + BN2 + PReLU2 fold
  size_t p_w2;                             // fp32: [9][mid][growth] ; bf16: [9][growth][mid]
  size_t p_b2;                             // [growth] fp32
};

struct BlockPlan {
  int H, W, Hp, Wp, R;  // interior size, padded size, rows per image (Hp*Wp)
  int c0, c0p, ctot;    // logical / padded input channels, physical channels of the block buffer
  int ld;               // row pitch of the block buffer in elements.  Eval walk: ctot rounded up to 64 channels, so that a
                        // row is a whole number of 128-byte lines and every 64-channel TMA box row is exactly ONE line.
                        // ncu on block 1 (ctot = 160, 320-byte rows): a 128-byte box row straddled two lines on every
                        // second row and DRAM fills whole lines - conv1 read 192 B per row for K = 64 and the whole
                        // 320 B for K = 96 / 128.  The pad channels are never written and never read (the tensor maps
                        // clip at ctot).  Training walk: ld == ctot (its passes read whole rows).
  int clog;             // logical output channels (c0 + growth*layers)
  std::vector<LayerPlan> layers;
  // transition after this block (absent after the last one)
  bool has_transition;
  BnArena tnorm;
  int64_t tconv_w, tconv_b;
  int tout, toutp;  // logical / padded output channels (= next block's c0 / c0p)
  int tkpad;        // ctot rounded up to kKChunk
  size_t p_t_scale, p_t_shift, p_t_alpha;  // [ctot] fp32
  size_t p_tw;                             // [ctot][toutp] fp32 (CUDA-core GEMM in both precisions)
  size_t p_tb;                             // [toutp] fp32
  int tn_tiles;                            // bf16 path: N tiles of 128 output channels
  size_t p_tw16;                           // bf16 [tn_tiles*128][tkpad], K-major, zero rows beyond tout
  size_t p_tb16, p_ta16;                   // [tn_tiles*128] fp32: bias (0-padded), PReLU slope 1 (identity)
  // workspace
  int chunk;      // images of this block processed per pass
  size_t ws_blk;  // byte offset
};

struct CnnPlan {
  tcvn_cnn_desc d;
  tcvn_precision prec;
  int esize;  // bytes per activation element
  int mid;    // bottleneck width
  int Hs, Ws;  // stem conv output (200 x 140)
  // arena
  int64_t conv0_w, conv0_b;
  BnArena norm0;
  BnArena final_norm;
  int64_t lin_w;
  BnArena out_norm;
  int64_t arena_floats;
  // packed
  size_t p_w0;                             // [in*49][init] fp32
  size_t p_s_scale, p_s_shift, p_s_alpha;  // [init] fp32: conv0 bias + BN0 + PReLU0
  size_t p_f_scale, p_f_shift, p_f_alpha;  // [ctot_last] fp32
  size_t p_lw;                             // [ctot_last][out] fp32
  size_t p_lo_scale, p_lo_shift, p_lo_alpha;  // [out] fp32
  size_t packed_bytes;
  std::vector<BlockPlan> blocks;
  // workspace (sized for `chunk` images)
  int chunk;
  size_t ws_stem, ws_mid, ws_pool, ws_gap, ws_hitofs, ws_bytes;

  static bool build(const tcvn_cnn_desc& d, tcvn_precision prec, int n_images, CnnPlan* out, bool pad_pitch = false);
  // physical channel of logical channel c inside block b's buffer
  int phys(int b, int c) const {
    const BlockPlan& B = blocks[b];
    return c < B.c0 ? c : c + (B.c0p - B.c0);
  }
};

inline bool CnnPlan::build(const tcvn_cnn_desc& d, tcvn_precision prec, int n_images, CnnPlan* out, bool pad_pitch) {
  if (d.num_blocks < 1 || d.num_blocks > TCVN_MAX_BLOCKS) return false;
  if (d.in_channels < 1 || d.init_features < 1 || d.growth < 1 || d.bn_size < 1 || d.out_features < 1) return false;
  static const bool pad_on = [] { const char* v = getenv("TCVN_PAD_PITCH"); return !(v && v[0] == '0'); }();   // A/B switch
  CnnPlan& P = *out;
  P.d = d;
  P.prec = prec;
  P.esize = prec == TCVN_BF16 ? 2 : 4;
  P.mid = d.bn_size * d.growth;
  P.Hs = (d.height + 6 - 7) / 2 + 1;
  P.Ws = (d.width + 6 - 7) / 2 + 1;
  int64_t a = 0;
  size_t p = 0;
  auto take = [&](int64_t n) { int64_t o = a; a += n; return o; };
  auto ptake = [&](size_t bytes) { size_t o = p; p += (bytes + 255) / 256 * 256; return o; };
  auto bn = [&](int c) {
    BnArena r;
    r.c = c;
    r.w = take(c); r.b = take(c); r.rm = take(c); r.rv = take(c);
    r.alpha = take(c);  // the PReLU weight follows its BatchNorm in state_dict order
    return r;
  };
  const size_t wsz = prec == TCVN_BF16 ? 2 : 4;
  P.conv0_w = take((int64_t)d.init_features * d.in_channels * 49);
  P.conv0_b = take(d.init_features);
  P.norm0 = bn(d.init_features);
  P.p_w0 = ptake((size_t)d.in_channels * 49 * d.init_features * 4);
  P.p_s_scale = ptake(d.init_features * 4);
  P.p_s_shift = ptake(d.init_features * 4);
  P.p_s_alpha = ptake(d.init_features * 4);
  int H = (P.Hs - 3) / 2 + 1, W = (P.Ws - 3) / 2 + 1;  // AvgPool2d(3, stride 2, no padding)
  int c = d.init_features;
  P.blocks.clear();
  for (int b = 0; b < d.num_blocks; ++b) {
    BlockPlan B;
    B.H = H; B.W = W; B.Hp = H + 2; B.Wp = W + 2; B.R = B.Hp * B.Wp;
    B.c0 = c;
    B.c0p = round_up(c, kChanAlign);
    const int nl = d.block_layers[b];
    if (nl < 0) return false;
    B.ctot = B.c0p + nl * d.growth;
    B.ld = pad_pitch && pad_on ? round_up(B.ctot, kKChunk) : B.ctot;
    B.clog = c + nl * d.growth;
    for (int i = 0; i < nl; ++i) {
      LayerPlan L;
      L.cin = c + i * d.growth;
      L.kphys = B.c0p + i * d.growth;
      L.kpad = round_up(L.kphys, kKChunk);
      L.norm1 = bn(L.cin);
      L.conv1_w = take((int64_t)P.mid * L.cin);
      L.conv1_b = take(P.mid);
      L.norm2 = bn(P.mid);
      L.conv2_w = take((int64_t)d.growth * P.mid * 9);
      L.conv2_b = take(d.growth);
      L.p_a_scale = ptake(L.kpad * 4);
      L.p_a_shift = ptake(L.kpad * 4);
      L.p_a_alpha = ptake(L.kpad * 4);
      L.p_w1 = ptake(prec == TCVN_BF16 ? (size_t)P.mid * L.kpad * wsz : (size_t)L.kphys * P.mid * wsz);
      L.p_o_scale = ptake(P.mid * 4);
      L.p_o_shift = ptake(P.mid * 4);
      L.p_o_alpha = ptake(P.mid * 4);
      L.p_w2 = ptake((size_t)9 * P.mid * d.growth * wsz);
      L.p_b2 = ptake(d.growth * 4);
      B.layers.push_back(L);
    }
    c = B.clog;
    B.has_transition = b != d.num_blocks - 1;
    if (B.has_transition) {
      B.tnorm = bn(c);
      B.tout = c / 2;
      B.toutp = round_up(B.tout, kChanAlign);
      B.tkpad = round_up(B.ctot, kKChunk);
      B.tconv_w = take((int64_t)B.tout * c);
      B.tconv_b = take(B.tout);
      B.p_t_scale = ptake(B.tkpad * 4);
      B.p_t_shift = ptake(B.tkpad * 4);
      B.p_t_alpha = ptake(B.tkpad * 4);
      B.p_tw = ptake((size_t)B.ctot * B.toutp * 4);
      B.p_tb = ptake(B.toutp * 4);
      B.tn_tiles = (B.toutp + 127) / 128;
      B.p_tw16 = ptake(prec == TCVN_BF16 ? (size_t)B.tn_tiles * 128 * B.tkpad * 2 : 0);
      B.p_tb16 = ptake(prec == TCVN_BF16 ? (size_t)B.tn_tiles * 128 * 4 : 0);
      B.p_ta16 = ptake(prec == TCVN_BF16 ? (size_t)B.tn_tiles * 128 * 4 : 0);
      c = B.tout;
      H /= 2; W /= 2;  // AvgPool2d(2, 2) floors
      if (H < 1 || W < 1) return false;
    }
    P.blocks.push_back(B);
  }
  const BlockPlan& last = P.blocks.back();
  P.final_norm = bn(c);
  P.lin_w = take((int64_t)d.out_features * c);
  P.out_norm = bn(d.out_features);
  P.arena_floats = a;
  P.p_f_scale = ptake(last.ctot * 4);
  P.p_f_shift = ptake(last.ctot * 4);
  P.p_f_alpha = ptake(last.ctot * 4);
  P.p_lw = ptake((size_t)last.ctot * d.out_features * 4);
  P.p_lo_scale = ptake(d.out_features * 4);
  P.p_lo_shift = ptake(d.out_features * 4);
  P.p_lo_alpha = ptake(d.out_features * 4);
  P.packed_bytes = p;
  // workspace: block b is processed `chunk` images at a time; later blocks have smaller maps and take whole
  // multiples of the previous block's chunk so that their launches still fill the 148 SMs.  The chunk is sized
  // by a working-set budget (concat buffer + bottleneck intermediate).  Measured on B200 (scripts/sweep_budget.sh,
  // 256-event batch; re-measured after conv1 became HBM-bound: 64 MB 6.1k, 128 MB 7.9k, 256 MB 9.1k, 768 MB 10.2k,
  // 1536 MB 10.6k, 3072 MB 10.5k): 48 MB 5.3k, 80 MB 6.6k, 160 MB 7.7k, 320 MB 8.7k, 768 MB 9.3k, 1024 MB 9.3k events/s - the
  // layer kernels are issue/latency-bound, not HBM-bound, so keeping a chunk inside the 126 MB L2 buys nothing
  // yet and long persistent launches win; revisit when the kernels approach the memory roofline.
  size_t l2_budget = (size_t)6144 << 20;   // re-measured after conv1 became bandwidth-bound: 768 MB 20.0-20.3 ms, 1536 MB 19.65-19.86, 3072 MB 19.6-19.9, 6144 MB 19.46-19.5 per 256-event step
  if (const char* e = getenv("TCVN_L2_BUDGET_MB")) { const int mb = atoi(e); if (mb > 0) l2_budget = (size_t)mb << 20; }
  // tiles of the two persistent kernels of a dense layer (128 / 126 rows) over `c` images: fraction of the
  // last wave of 148 CTAs that does useful work
  auto wave_eff = [&](const BlockPlan& B, int c) {
    const long long rows = (long long)c * B.R;
    const long long t1 = (rows + 127) / 128, t2 = (rows + 125) / 126;
    const double e1 = (double)t1 / (double)((t1 + 147) / 148 * 148), e2 = (double)t2 / (double)((t2 + 147) / 148 * 148);
    return e1 < e2 ? e1 : e2;
  };
  int prev = 0;
  for (auto& B : P.blocks) {
    const size_t per_image = (size_t)B.R * (B.ld + P.mid) * P.esize;
    int c = (int)(l2_budget / per_image);
    if (c < 1) c = 1;
    if (prev) c = c / prev * prev;
    if (c < prev) c = prev;
    if (c > 4096) c = prev ? 4096 / prev * prev : 4096;
    if (prev == 0) {
      if (c > n_images) c = n_images < 1 ? 1 : n_images;
      else {
        // first block: among chunk sizes within 25 % of the budget pick the one that fills its last wave best
        int best = c;
        double best_eff = wave_eff(B, c);
        for (int k = c - 1; k >= c - c / 4 && k >= 1; --k) {
          const double e = wave_eff(B, k);
          if (e > best_eff + 0.02) { best = k; best_eff = e; }
        }
        c = best;
      }
      if (const char* e = getenv("TCVN_CHUNK0")) { const int f = atoi(e); if (f > 0 && f <= n_images) c = f; }
    }
    if (prev && c > n_images) c = (n_images + prev - 1) / prev * prev;  // never larger than the batch needs
    B.chunk = c;
    prev = c;
  }
  P.chunk = P.blocks[0].chunk;
  size_t w = 0;
  auto wtake = [&](size_t bytes) { size_t o = w; w += (bytes + 1023) / 1024 * 1024; return o; };
  P.ws_stem = 0;  // the stem is fused with its pooling: no pre-pool map in memory
  size_t mid_rows = 0, pool_elems = 0;
  for (auto& B : P.blocks) {
    B.ws_blk = wtake((size_t)B.chunk * B.R * (size_t)B.ld * P.esize);
    if ((size_t)B.chunk * B.R > mid_rows) mid_rows = (size_t)B.chunk * B.R;
  }
  for (size_t b = 0; b + 1 < P.blocks.size(); ++b) {
    size_t e = (size_t)P.blocks[b].chunk * P.blocks[b + 1].R * P.blocks[b].ctot;
    if (e > pool_elems) pool_elems = e;
  }
  P.ws_mid = wtake(mid_rows * P.mid * P.esize);
  P.ws_pool = wtake(pool_elems * P.esize + 1024);
  P.ws_gap = wtake((size_t)P.blocks.back().chunk * (size_t)last.ctot * 4);
  P.ws_hitofs = wtake(((size_t)(n_images < 1 ? 1 : n_images) + 1) * sizeof(long long));
  P.ws_bytes = w;
  return true;
}

}  // namespace tcvn
