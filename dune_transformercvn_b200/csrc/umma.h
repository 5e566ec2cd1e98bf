// bf16 tcgen05 path: launch wrappers used by the CNN walker.
#pragma once
#include <cuda.h>

#include "kernels.h"

namespace tcvn {
int umma_dense_layer(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* packed, void* blk, void* mid,
                     long long rows, cudaStream_t st);
int umma_dense_layer_part(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* packed, void* blk,
                          void* mid, long long rows, int which, cudaStream_t st);
int umma_transition(const CnnPlan& P, const BlockPlan& B, const BlockPlan& Nx, const char* packed, const void* pool,
                    void* next_blk, long long rows, cudaStream_t st);

// shared with the training kernels (umma_train.cu)
int make_map(const void* base, long long rows, int cols, int pitch, int box_cols, int box_rows, CUtensorMap* out);
int sm_count();
bool& device_flag(int site);
// out[m, n] = PReLU_n( sum_k act_k(A[m, k]) * W[n, k] + shift[n] ), bf16 in / bf16 out, N tiles of 128 (see umma.cu)
int launch_gemm(bool transform, const void* A, long long rows, int a_cols, int a_pitch, const void* W, int w_rows, int kpad,
                int kphys, const float* a_scale, const float* a_shift, const float* a_alpha, const float* o_shift,
                const float* o_alpha, void* out, int out_cols, int out_pitch, int n_tiles_n, int Hp, int Wp, cudaStream_t st,
                double* stats = nullptr, int* stat_slots = nullptr);
int launch_gemm_shifted(const void* A, long long rows, int a_cols, int n_taps, const int* tap_off, const void* X2, int x2_cols,
                        const void* W, int n_tiles_n, const float* bias, const float* ones, void* out, int out_cols, int Hp, int Wp,
                        cudaStream_t st, int a_pitch = 0, const int* tap_col = nullptr);
// a_pitch > a_cols with tap_col: every tap reads its own a_cols-wide column group of a wider A (space-to-depth input of a
// stride-2 convolution: tap (dy, dx) = parity group (dy & 1, dx & 1) shifted by (dy >> 1, dx >> 1))
// slots of the epilogue statistics (doubles [slots][2][128] for launch_gemm, [slots][2][32] for umma_conv2_fwd)
constexpr int kUmmaStatSlotsMax = 8 * 148;
int umma_conv2_fwd(const void* mid, long long rows, const void* w2, const float* bias, void* out, int ldo, int col0, int Hp,
                   int Wp, int W, cudaStream_t st, float p_drop = 0.f, unsigned long long seed = 0, unsigned long long site = 0,
                   double* stats = nullptr, int* stat_slots = nullptr);
size_t umma_wgrad_parts_bytes(int n_items);
int reduce_parts(const float* parts, int n_parts, long long n, float* out, bool accumulate, cudaStream_t st);
int umma_wgrad(const void* A, long long rows, int a_cols, int a_pitch, int n_items, const int* item_col, const int* item_shift,
               const int* item_valid, const float* a_scale, const float* a_shift, const float* a_alpha, int a_fold_cols,
               const void* G, int g_cols, int g_pitch, int g_col0, float* parts, float* dw, bool accumulate, cudaStream_t st,
               int map_mode = 0, int map_n_log = 0, int map_k_log = 0, int map_c0 = 0, int map_c0p = 0);
// map_mode != 0: dw is the reference-layout gradient tensor and the per-CTA partial sums are ADDED to it re-laid out in the
// same launch (1: 1x1 convolution dst[n][k] over logical channels; 2: 3x3 convolution dst[n][k][dy][dx]); see umma_train.cu
// red_x != null: the BN2 + PReLU2 backward reductions are fused into the epilogue (see DgradParams): red_parts[slots][3][128]
int umma_conv2_dgrad(const void* g2x, const void* wd, long long rows, int Hp, int Wp, void* out, cudaStream_t st,
                     const void* red_x = nullptr, const float* red_fold = nullptr, double* red_parts = nullptr,
                     int* red_slots = nullptr);
}  // namespace tcvn
