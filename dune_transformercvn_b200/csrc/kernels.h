// Internal launch wrappers shared by the CNN walker (cnn.cu).  Not part of the C ABI.
#pragma once
#include "common.cuh"
#include "plan.h"

#include <vector>

namespace tcvn {

// out[m, col0 + n] = epi( sum_t sum_k act(A[m + tap_off[t], k]) * W[t][k][n] )   (fp32 CUDA-core path)
//   act(x)  = prelu(x * a_scale[k] + a_shift[k], a_alpha[k])   when a_scale != nullptr, else x
//   epi(v)  = prelu(v * o_scale[n] + o_shift[n], o_alpha[n])   when o_scale != nullptr, else v + o_shift[n]
//   rows of the zero ring (ring_R > 0) are written as zeros; rows outside [0, m_total) read as zeros.
struct GemmArgs {
  const void* A; int lda; long long m_total; int K; int taps; int tap_off[9];
  const float* W; int N;
  const float* a_scale; const float* a_shift; const float* a_alpha;
  const float* o_scale; const float* o_shift; const float* o_alpha;
  void* out; int ldo; int out_col0;
  int ring_Hp, ring_Wp;  // 0 = no ring masking
  bool a_is_f32; bool out_is_f32;  // element types of A / out (false = bf16)
  int a_ring_Hp = 0, a_ring_Wp = 0;  // > 0: activated A rows on the zero ring read as 0
  bool accumulate = false;           // out += result
};
int launch_simt_gemm(const GemmArgs& g, cudaStream_t stream);

// stem, fused: conv0 7x7 s2 p3 + bias + BN0 + PReLU0 + AvgPool2d(3,2) on NCHW fp32 pixels -> channels [0, c0) of
// block 0's ringed buffer [n, Hb+2, Wb+2, ldo]
int launch_stem(const float* pixels, int n, int cin, int H, int W, const float* w0, const float* s_scale,
                const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb, int Wb, bool f32,
                cudaStream_t stream);
// stem straight from the COO hit list (no dense map): image_offsets[i] = first hit of image i
int launch_hit_offsets(const int32_t* coords, long long nnz, int n_images, long long* offsets, cudaStream_t stream);
int launch_stem_coo(const int32_t* coords, const void* values, bool values_u8, const long long* image_offsets, int image0,
                    float divisor, int n, int cin, int H, int W, const float* w0, const float* s_scale,
                    const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb, int Wb, bool f32,
                    cudaStream_t stream);
// binned variant (stem_coo.cu): the hits of ALL images of a call are binned once by 4 x 4-pooled-pixel tile
// (order-preserving), then every block-0 chunk runs the warp-per-tile kernel over its images
size_t stem_bins_bytes(int n_images, int Hb, int Wb, long long nnz);
int launch_stem_bin(const int32_t* coords, const void* values, bool values_u8, const long long* image_offsets, int n_images,
                    long long nnz, int cin, float divisor, int H, int W, int Hb, int Wb, void* bins, cudaStream_t stream);
int launch_stem_coo_binned(int image0, int n, int n_images_binned, long long nnz_binned, int cin, int H, int W, const float* w0,
                           const float* s_scale, const float* s_shift, const float* s_alpha, int c0, void* blk, int ldo, int Hb,
                           int Wb, bool f32, const void* bins, cudaStream_t stream);
// transition front half: BN + PReLU + AvgPool2d(2,2) of a ringed block buffer into a ringed buffer of the next geometry
int launch_act_pool2(const void* blk, int n, int H, int W, int ld, int c, const float* scale, const float* shift,
                     const float* alpha, void* out, int H2, int W2, bool f32, cudaStream_t stream);
// tail front half: BN + PReLU + global average over the interior -> gap [n, c] fp32
int launch_act_gap(const void* blk, int n, int H, int W, int ld, int c, const float* scale, const float* shift,
                   const float* alpha, float* gap, bool f32, cudaStream_t stream);
// ringed channels-last -> NCHW fp32 (test hook)
int launch_ring_to_nchw(const void* blk, int n, int H, int W, int ld, int c, int c_skip_from, int c_skip, float* out,
                        bool f32, cudaStream_t stream);

// pack.cu helpers (also used by seq.cu)
int fold(const float* arena, const BnArena& bn, const float* conv_bias, int c0, int c0p, int n_out, float eps,
         char* packed, size_t o_scale, size_t o_shift, size_t o_alpha, cudaStream_t st);
int repack(const float* src, int n_log, int k_log, int taps, int c0, int c0p, int K_out, int N_out, bool k_major,
           bool bf16, void* dst, cudaStream_t st, const float* row_bn_w = nullptr, const float* row_bn_rv = nullptr,
           float eps = 0.f);
int pad_copy(const float* src, int n_log, float* dst, int n_out, cudaStream_t st);
// a list of re-layouts executed in one launch per 96 entries (pack.cu).  k_major: 0 dst[tap][k][n], 1 dst[tap][n][k],
// 2 conv2 weight -> Wd[dy][c][dx*32+n] bf16 (the conv2 input-gradient operand; the shape arguments are ignored
// except taps*K_out*N_out = 3*128*128)
struct RepackList {
  struct Entry { unsigned char raw[80]; };
  std::vector<Entry> items;
  void add(const float* src, int n_log, int k_log, int taps, int c0, int c0p, int K_out, int N_out, int k_major, bool bf16,
           void* dst);
  int run(cudaStream_t st);
};

// train.cu: typed cores of the training primitives (bool *_bf16 = element type of that operand; false = fp32).
// Column reductions are bit-reproducible: per-slab partial sums in fixed slots, added in a fixed order.
size_t colsum_parts_bytes();   // scratch for the partial sums of one reduction
int colsums_typed(int mode, const void* X, bool x_bf16, int ldx, int xcol0, const void* D, bool d_bf16, int ldd, int dcol0,
                  const float* fold, int fold_stride, int C, long long m_total, int ring_hp, int ring_wp, double* out,
                  int out_stride, double* parts, cudaStream_t stream);
int colsums_into(int mode, const float* X, int ldx, int xcol0, const float* D, int ldd, int dcol0, const float* fold, int C,
                 long long m_total, int ring_hp, int ring_wp, double* out, int out_stride, double* parts, cudaStream_t stream);
// partial sums only: parts[slab][ns][C], ns = 2 / 3 / 1 for mode 0 / 1 / 2; the consumer adds the slabs
int colsums_parts(int mode, const void* X, bool x_bf16, int ldx, int xcol0, const void* D, bool d_bf16, int ldd, int dcol0,
                  const float* fold, int fold_stride, int C, long long m_total, int ring_hp, int ring_wp, double* parts,
                  int* n_slabs, cudaStream_t stream);
int bnact_bwd_apply_typed(const void* D, bool d_bf16, int ldd, int dcol0, const void* X, bool x_bf16, int ldx, int xcol0,
                          const float* fold, int fold_stride, const double* sums, int C, double count, void* dX, bool o_bf16,
                          int lddx, int dxcol0, bool accumulate, long long m_total, int ring_hp, int ring_wp,
                          cudaStream_t stream);
int bnact_fwd_typed(const void* X, bool x_bf16, int ldx, int xcol0, const float* fold, int fold_stride, int C, long long m_total,
                    int ring_hp, int ring_wp, void* out, bool o_bf16, int ldo, int ocol0, cudaStream_t stream);
int pool_typed(int kind, const void* src, const float* fold, void* dst, bool bf16, int n, int C, int H, int W, int H2, int W2,
               int ld, cudaStream_t stream);
int stem_conv_typed(const float* pixels, int n, int cin, int H, int W, const float* w, const float* bias, int C, float* z,
                    const void* dz, bool dz_bf16, float* dw, cudaStream_t stream);
int dropout_typed(void* X, bool bf16, int ld, int col0, int C, long long m_total, uint64_t seed, uint64_t stream_id, float p,
                  cudaStream_t stream);
// BN1 + PReLU1 backward reductions of a dense layer (train.cu): parts[slab][3][C]
int bn1_bwd_reduce(const void* X, int ldx, const void* D, int ldd, const float* fold, int fold_stride, int C, long long m_total,
                   int ring_hp, int ring_wp, double* parts, int* n_slabs, cudaStream_t stream);
// train_stem.cu: the stem of the training path, hit-driven and bit-reproducible
int stem_train_slots(int n_images, int H, int W);
int stem_train_forward(const float* pixels, int n, int cin, int H, int W, const float* w0, const float* bias, int C, void* z0_bf16,
                       double* stat_parts, int* n_slots, cudaStream_t stream);
int stem_train_wgrad(const float* pixels, int n, int cin, int H, int W, const void* dz_bf16, int C, float* dw_parts, int* n_slots,
                     cudaStream_t stream);
// AvgPool2d(3, 2) of act(z0) into block 0 (forward) and its backward, all-bf16 maps, 8 channels per thread (train.cu)
int stem_pool16_forward(const void* z0_bf16, const float* fold, void* blk_bf16, int n, int C, int H, int W, int Hs, int Ws, int ld,
                        cudaStream_t stream);
int stem_pool16_backward(const void* dblk_bf16, int ld, void* dz_bf16, int n, int C, int H, int W, int Hs, int Ws, cudaStream_t stream);
int wgrad_typed(const void* A, bool a_bf16, int lda, long long m_total, int K, int taps, const int* tap_off, const float* a_scale,
                const float* a_shift, const float* a_alpha, int a_ring_hp, int a_ring_wp, const void* G, bool g_bf16, int ldg,
                int g_col0, int N, int g_ring_hp, int g_ring_wp, float* dW, cudaStream_t stream, float* parts = nullptr,
                size_t parts_floats = 0);
// dW[K][N] += A^T G over m_total rows (plain fp32 matrices); parts = scratch for per-slab partial results (nullable)
int wgrad_f32(const float* A, int lda, long long m_total, int K, const float* G, int ldg, int N, float* dW, float* parts,
              size_t parts_floats, cudaStream_t stream);

#define TCVN_TRY(expr)                \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != TCVN_OK) return rc__; \
  } while (0)

}  // namespace tcvn
