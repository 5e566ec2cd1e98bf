/* tcvn.h — C ABI of the B200-native TransformerCVN hot path (libtcvn.so, sm_100a only).
 *
 * The reference (ayankele/dune-transformercvn) has no FFI: its "plugin API" for this path is
 * Python attribute lookup.  Each entry point below names the reference
 * interface it stands in for (paths relative to the reference tree).  The Python host side
 * (dune_transformercvn_b200/) binds these with ctypes; INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative tcvn_status; tcvn_last_error() gives
 *     a thread-local message.  Nothing throws, nothing allocates device memory, nothing
 *     synchronises the stream: the caller owns every buffer and the stream.  Host-side state is
 *     per calling thread only: the error text, the seed-offset / SM-budget settings below and a
 *     bounded memo of encoded TMA tensor maps (csrc/umma.cu: make_map); there are no locks.
 *   - all pointers except descriptors are DEVICE pointers unless the name ends in _host.
 *   - workspace sizes are queried with the matching *_bytes function.
 *   - "arena" = the fp32 tensors of a reference sub-module's state_dict, concatenated in
 *     state_dict order, int64 num_batches_tracked counters omitted
 *     (dune_transformercvn_b200/params.py pins that order against the reference).
 *   - activations are channels-last with a one-pixel zero ring: a feature map of N images,
 *     H x W pixels and C channels is a row-major matrix [N*(H+2)*(W+2), C]; a 3x3 tap is then
 *     a constant row offset, and the DenseNet concat is a column slice written in place.
 */
#ifndef TCVN_H_
#define TCVN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCVN_ABI_VERSION 2
#define TCVN_MAX_BLOCKS 8
#define TCVN_MAX_DECODER_LAYERS 8

typedef struct CUstream_st* tcvn_stream_t;

typedef enum tcvn_status {
  TCVN_OK = 0,
  TCVN_ERR_ARG = -1,        /* bad descriptor / null pointer / size mismatch            */
  TCVN_ERR_WORKSPACE = -2,  /* caller's workspace or packed buffer is too small         */
  TCVN_ERR_CUDA = -3,       /* a CUDA runtime / driver call failed (launch, tensor map) */
  TCVN_ERR_UNSUPPORTED = -4 /* configuration outside what the kernels were written for  */
} tcvn_status;

typedef enum tcvn_precision {
  TCVN_FP32 = 0, /* fp32 storage, fp32 FMA accumulation (CUDA cores): the 1e-4 parity path   */
  TCVN_BF16 = 1  /* bf16 storage, tcgen05 MMA with fp32 TMEM accumulators: the throughput path */
} tcvn_precision;

typedef enum tcvn_value_dtype { TCVN_VAL_F32 = 0, TCVN_VAL_U8 = 1 } tcvn_value_dtype;

typedef enum tcvn_dense_layout {
  TCVN_NCHW_F32 = 0 /* what transformercvn's sparse_to_dense returns */
} tcvn_dense_layout;

int tcvn_abi_version(void);
const char* tcvn_last_error(void);
/* kernels launched by this library in this process so far (bench.py reports the per-step count) */
long long tcvn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1  ingest: Minkowski-format COO hits -> dense pixel maps, value / divisor fused.
 * Replaces: sparse_to_dense(features, coordinates, image_size)
 *           transformercvn/network/trainers/neutrino_full_dense_trainer.py:15-24
 *           and the "/ 255.0" of preprocess_pixels, same file :59-60.
 * coords (nnz,3) int32 rows [image, y, x] sorted by image; values (nnz,channels) f32 or u8;
 * out (n_images, channels, height, width) fp32, fully written (zeros + hits) in one pass.
 * divisor == 0 means "no scaling".  Bit-exact against the reference (IEEE fp32 division).   */
int tcvn_densify(const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                 int channels, int n_images, int height, int width, float divisor, float* out,
                 tcvn_dense_layout layout, tcvn_stream_t stream);
/* Same with the training-time pixel noise of preprocess_pixels fused in (same file :62-65: values *= 1 + randn * std,
 * std = options.pixel_noise_std): one standard-normal draw per stored value from a counter hash of (seed, hit, channel),
 * so a (seed, input) pair always gives the same map.  noise_std == 0 is bit-identical to tcvn_densify.
 * Hits must be sorted by image (the reference's index_put does not need that; its file format guarantees it). */
int tcvn_densify_noise(const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                       int channels, int n_images, int height, int width, float divisor, float noise_std, uint64_t seed,
                       float* out, tcvn_dense_layout layout, tcvn_stream_t stream);

/* Collate: the events' COO hit lists concatenated in event order, image index still LOCAL to the event -> batch-global
 * image indices (what tcvn_densify / tcvn_cnn_forward_sparse consume).
 * Replaces: MinkowskiCollection.collate_sparse  transformercvn/dataset/minkowski_dataset.py:34-47 (Python loop,
 *           coord[:,0] += images of all earlier events, one .item() sync per event).
 * hits_per_event (n_events) int32; the number of images of an event is images_per_event[e] or, when that is NULL,
 * the number of set flags in masks[e, 0..max_slots) (the reference uses mask.sum()).  coords_out may alias coords_in. */
size_t tcvn_collate_workspace_bytes(int n_events);
int tcvn_collate_coords(const int32_t* coords_in, int64_t nnz, const int32_t* hits_per_event, const int32_t* images_per_event,
                        const uint8_t* masks, int n_events, int max_slots, int32_t* coords_out, void* workspace,
                        size_t workspace_bytes, tcvn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DenseNet pixel-map embedding.
 * Replaces: transformercvn/network/layers/dense_net.py:97-167 (DenseNet), :8-45 (Bottleneck),
 *           :78-94 (Transition), constructed by
 *           transformercvn/network/networks/neutrino_full_dense_network.py:6-16.            */
typedef struct tcvn_cnn_desc {
  int32_t in_channels;   /* 3                      dense_net.py:112                      */
  int32_t init_features; /* 64  initial_pixel_dim  dense_net.py:112-118                  */
  int32_t growth;        /* 32  densenet_growth_rate                                     */
  int32_t bn_size;       /* 4   bottleneck width = bn_size * growth                      */
  int32_t num_blocks;    /* 5                                                            */
  int32_t block_layers[TCVN_MAX_BLOCKS]; /* 3,6,12,6,3                                  */
  int32_t out_features;  /* 256 (prong CNN) / 288 (event CNN)                            */
  int32_t height, width; /* 400, 280                                                     */
  float bn_eps;          /* 1e-5 (torch default)                                         */
} tcvn_cnn_desc;

/* number of fp32 values in the CNN's arena (state_dict order, counters omitted) */
int64_t tcvn_cnn_arena_floats(const tcvn_cnn_desc* d);
/* device bytes of the packed (kernel-ready) parameter block for inference */
size_t tcvn_cnn_packed_bytes(const tcvn_cnn_desc* d, tcvn_precision prec);
/* arena -> packed: folds eval-mode BatchNorm (running stats) into per-channel scale/shift,
 * re-lays conv weights K-major with the channel padding of the in-place concat buffers,
 * converts to bf16 for TCVN_BF16.  Runs on the stream; call again after the weights change. */
int tcvn_cnn_pack(const tcvn_cnn_desc* d, tcvn_precision prec, const float* arena, void* packed,
                  size_t packed_bytes, tcvn_stream_t stream);
/* workspace for a forward pass over n_images images.  Layout (csrc/plan.h): one ringed NHWC buffer per dense block, rows
 * padded to whole 128-byte lines (channel total rounded up to 64: a 64-channel TMA box row is then exactly one line; the
 * pad channels are never written or read), the 128-channel bottleneck map, pooling scratch.  Images are walked in chunks
 * sized by a working-set budget (6 GB by default, TCVN_L2_BUDGET_MB), so the size grows with n_images only up to the chunk. */
size_t tcvn_cnn_workspace_bytes(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images);
/* eval-mode forward: pixels (n_images, in_channels, height, width) fp32 NCHW  ->
 * embedding (n_images, out_features) fp32.  The workspace must have been zero-filled once
 * after allocation (its padding rings are never written).                                  */
int tcvn_cnn_forward(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed, const float* pixels,
                     int n_images, float* embedding, void* workspace, size_t workspace_bytes,
                     tcvn_stream_t stream);
/* workspace for tcvn_cnn_forward_sparse including the scratch of the binned stem (hits binned by stem tile, order
 * preserving; tiles without hits only store the per-channel constant).  With a workspace of only tcvn_cnn_workspace_bytes
 * the sparse forward falls back to the unbinned stem kernel (same bits, slower). */
size_t tcvn_cnn_workspace_bytes_sparse(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images, int64_t nnz);
/* Same forward fed straight from the Minkowski-format hit list: fuses sparse_to_dense and the
 * "/ divisor" of preprocess_pixels (trainers/neutrino_full_dense_trainer.py:15-24,59-60) into the stem,
 * so the dense 3 x H x W map is never built.  coords (nnz,3) int32 [image, y, x] sorted by image,
 * values (nnz, in_channels) f32 or u8, coordinates unique per image (as the dataset guarantees).     */
int tcvn_cnn_forward_sparse(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed,
                            const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                            float divisor, int n_images, float* embedding, void* workspace,
                            size_t workspace_bytes, tcvn_stream_t stream);
/* measurement hook (bench.py roofline): re-launches ONE kernel of dense layer `layer` of block `block` on the
 * feature maps a previous forward left in the workspace.  which: 1 = conv1 (fused BN+PReLU GEMM), 2 = conv2
 * (3x3 shifted GEMM), 3 = both.  TCVN_BF16 only.                                                          */
int tcvn_cnn_run_layer(const tcvn_cnn_desc* d, tcvn_precision prec, const void* packed, void* workspace,
                       size_t workspace_bytes, int n_images, int block, int layer, int which, tcvn_stream_t stream);
/* test hook: copies one internal feature map of the last forward (still in the workspace) to
 * out as (n_images, channels, h, w) fp32 NCHW.  stage: 0 = stem after pool, 2b+1 = dense block
 * b output, 2b+2 = transition b output (b from 0).  Returns the channel count in *channels.   */
int tcvn_cnn_read_stage(const tcvn_cnn_desc* d, tcvn_precision prec, const void* workspace, int n_images,
                        int stage, float* out, int32_t* channels, int32_t* h, int32_t* w,
                        tcvn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Token assembly + transformer encoder + classification heads (eval mode).
 * Replaces: BaseProngEmbedding.forward  networks/neutrino_full_base_network.py:87-125
 *           (with layers/packed_data.py:59-76 and layers/prong_feature_embedding.py:7-33),
 *           ProngCustomBertEncoder.forward  layers/prong_custom_bert_encoder.py:57-75,
 *           ProngDecoder  layers/prong_decoder.py:13-16,
 *           ProngTargetDecoder  layers/prong_target_decoder.py:35-41.                      */
typedef struct tcvn_seq_desc {
  int32_t hidden;        /* 128 */
  int32_t heads;         /* 8   */
  int32_t layers;        /* 6   */
  int32_t ffn;           /* 128 (dim_feedforward = hidden_dim, prong_custom_bert_encoder.py:45-52) */
  int32_t pixel_dim;     /* 256 */
  int32_t feature_dim;   /* 32  (zeros when disable_smart_features) */
  int32_t position_dim;  /* 32  */
  int32_t num_event_classes; /* 4 */
  int32_t num_prong_classes; /* 8 */
  int32_t num_decoder_layers;                    /* 4  */
  int32_t decoder_widths[TCVN_MAX_DECODER_LAYERS]; /* 64,32,16,8 */
  float bn_eps;          /* 1e-5 */
  float ln_eps;          /* 1e-5 */
} tcvn_seq_desc;

size_t tcvn_seq_packed_bytes(const tcvn_seq_desc* d);
/* Arenas (fp32, state_dict order): position = prong_embedding.event_position_embedding (1,P)
 * — the reference uses the EVENT vector for prong rows too (neutrino_full_base_network.py:107);
 * combined = prong_embedding.combined_embedding.*; encoder = encoder.encoder.layers.*;
 * event_decoder = event_decoder.hidden_layer.*; prong_decoder = prong_decoder.*             */
int tcvn_seq_pack(const tcvn_seq_desc* d, const float* position, const float* combined, const float* encoder,
                  const float* event_decoder, const float* prong_decoder, void* packed, size_t packed_bytes,
                  tcvn_stream_t stream);

#define TCVN_SEQ_TOKENS 1  /* embeddings -> tokens (B,S,hidden)                      */
#define TCVN_SEQ_ENCODER 2 /* tokens -> hidden (S,B,hidden), masked                  */
#define TCVN_SEQ_HEADS 4   /* hidden -> event_logits (B,E), prong_logits (B,L,C)     */
/* stages: OR of the flags above; a skipped earlier stage reads its output buffer as input.
 * event_embedding (B, pixel_dim+feature_dim), prong_embedding (T, pixel_dim) packed in
 * (event, slot) order of the set bits of prong_mask (B,L) (uint8/bool); event_mask (B,1) may be
 * NULL (= all true).  S = 1 + L <= 32.  tokens / hidden may be NULL when all three stages run. */
size_t tcvn_seq_workspace_bytes(const tcvn_seq_desc* d, int n_events, int max_prongs);
int tcvn_seq_forward(const tcvn_seq_desc* d, const void* packed, int stages, const float* event_embedding,
                     const float* prong_embedding, const uint8_t* event_mask, const uint8_t* prong_mask,
                     int n_events, int max_prongs, float* tokens, float* hidden, float* event_logits,
                     float* prong_logits, void* workspace, size_t workspace_bytes, tcvn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training primitives (fp32): the kernels behind the train-mode forward and the hand-written backward.
 * Row-matrix convention: a feature map is the ringed channels-last matrix (ring_hp = H+2, ring_wp = W+2; ring
 * rows carry neither data nor gradient), token / image level tensors are plain [rows, C] (ring_hp = 0).
 * fold = [scale | shift | alpha | mean | rstd] (5 x C) of one BatchNorm+PReLU pair for the current batch.
 * Replaces autograd of: nn.Conv2d / nn.Linear (dense_net.py:21-38,87-92,158), nn.BatchNorm2d/1d in train mode,
 * nn.PReLU, nn.AvgPool2d, nn.Dropout, nn.TransformerEncoderLayer (prong_custom_bert_encoder.py:45-54).       */
/* out[m, col0+n] (+)= bias[n] + sum_t sum_k act(A[m+tap_off[t], k]) * W[t][k][n];  act = PReLU(BN) from a_fold or identity */
int tcvn_t_gemm(const float* A, int lda, int64_t m_total, int K, int taps, const int32_t* tap_off, const float* W, int N,
                const float* a_fold, int a_ring_hp, int a_ring_wp, const float* bias, float* out, int ldo, int out_col0,
                int out_ring_hp, int out_ring_wp, int accumulate, tcvn_stream_t stream);
/* dW[t][k][n] += sum_m act(A[m+tap_off[t], k]) * G[m, g_col0+n]   (atomic accumulation; caller zeroes dW) */
int tcvn_t_wgrad(const float* A, int lda, int64_t m_total, int K, int taps, const int32_t* tap_off, const float* a_fold,
                 int a_ring_hp, int a_ring_wp, const float* G, int ldg, int g_col0, int N, int g_ring_hp, int g_ring_wp,
                 float* dW, tcvn_stream_t stream);
/* column sums over interior rows into doubles: mode 0 (sum x, sum x^2), mode 1 BN+PReLU backward reductions
 * (sum g, sum g*xhat, sum dA*min(y,0)), mode 2 (sum x) */
int tcvn_t_colsums(int mode, const float* X, int ldx, int xcol0, const float* D, int ldd, int dcol0, const float* fold, int C,
                   int64_t m_total, int ring_hp, int ring_wp, double* sums, tcvn_stream_t stream);
/* batch statistics -> fold; updates running_mean / running_var in place (momentum, unbiased variance) when given */
int tcvn_t_bn_finalize(const double* sums, int C, double count, const float* gamma, const float* beta, const float* alpha,
                       float eps, float momentum, float* running_mean, float* running_var, float* fold, tcvn_stream_t stream);
/* dX (+)= BN+PReLU backward of D at X; dgamma/dbeta/dalpha += the reductions (each nullable) */
int tcvn_t_bnact_bwd_apply(const float* D, int ldd, int dcol0, const float* X, int ldx, int xcol0, const float* fold,
                           const double* sums, int C, double count, float* dX, int lddx, int dxcol0, int accumulate,
                           int64_t m_total, int ring_hp, int ring_wp, float* dgamma, float* dbeta, float* dalpha,
                           tcvn_stream_t stream);
int tcvn_t_add_colsums(const double* sums, int C, float* dst, tcvn_stream_t stream);
int tcvn_t_bnact_fwd(const float* X, int ldx, int xcol0, const float* fold, int C, int64_t m_total, int ring_hp, int ring_wp,
                     float* out, int ldo, int ocol0, tcvn_stream_t stream);
/* kind 0 stem BN+PReLU+AvgPool(3,2) fwd, 1 its backward to the activated stem map, 2 AvgPool(2,2) backward, 3 global
 * average pool backward (see csrc/train.cu) */
int tcvn_t_pool(int kind, const float* src, const float* fold, float* dst, int n, int C, int H, int W, int H2, int W2, int ld,
                tcvn_stream_t stream);
int tcvn_t_dropout(float* X, int ld, int col0, int C, int64_t m_total, uint64_t seed, uint64_t stream_id, float p,
                   tcvn_stream_t stream);
/* raw conv0 (7x7 s2 p3) from NCHW pixels, hit-driven: forward (dz NULL) z = bias + conv; backward dw += x (*) dz */
int tcvn_t_stem_conv(const float* pixels, int n, int cin, int H, int W, const float* w, const float* bias, int C, float* z,
                     const float* dz, float* dw, tcvn_stream_t stream);
/* dir 0: out = LayerNorm(a + b), saves pre-norm sum and (mean, rstd); dir 1: a = dOut -> out = d(pre), dgamma/dbeta += */
int tcvn_t_layernorm(int dir, const float* a, const float* b, int D, int R, const float* gamma, const float* beta, float eps,
                     float* pre, float* stats, float* out, float* dgamma, float* dbeta, tcvn_stream_t stream);
/* dir 0: P = softmax(mask(QK^T/sqrt(dh))), ctx = dropout(P) V; dir 1: dqkv from dctx */
int tcvn_t_attention(int dir, const float* qkv, const uint8_t* mask, int B, int S, int heads, int D, float* P,
                     float* ctx_or_dctx, float* dqkv, float p_drop, uint64_t seed, uint64_t stream_id, tcvn_stream_t stream);
/* kind 0 gelu, 1 gelu backward (a = pre-activation, b = dOut), 2 a + b, 3 rows of a times mask */
int tcvn_t_eltwise(int kind, const float* a, const float* b, const uint8_t* mask, int C, int64_t total, float* out,
                   tcvn_stream_t stream);
/* token plumbing: op 0 offsets from mask, 1 token-input gather (dir 1: gradient back), 2 rows <-> padded sequence,
 * 3 prong-head rows */
int tcvn_t_tokens(int op, int dir, float* a, float* b, const float* pos, const uint8_t* prong_mask, int* offsets, int B, int T,
                  int L, int D, int pixel_dim, int feature_dim, int position_dim, float* x, tcvn_stream_t stream);
/* BN+PReLU+AvgPool(2,2) and BN+PReLU+global-average with a batch-statistics fold (fp32) */
int tcvn_t_act_pool2(const float* blk, int n, int H, int W, int ld, int C, const float* fold, float* out, int H2, int W2,
                     tcvn_stream_t stream);
int tcvn_t_act_gap(const float* blk, int n, int H, int W, int ld, int C, const float* fold, float* gap, tcvn_stream_t stream);
/* fused AdamW over the flat arenas: ONE step of every parameter group in one launch; replaces torch.optim.AdamW.step
 * (trainers/neutrino_base.py:109-130) and Lightning's clip_grad_norm_ (train.py:140).
 * select (one byte per element; may be NULL for a single group): select[i] = k > 0 -> element i belongs to group k-1 (the
 * reference's decay / no-decay split, neutrino_base.py:116-128); 0 marks buffers and parameters without a gradient.
 * lr / beta1 / beta2 / eps / weight_decay / step: HOST arrays of n_groups entries (step >= 1 = the group's step count).
 * max_norm > 0: the squared gradient norm is first reduced over the whole arena (one partial sum per block in the
 * workspace, added in a fixed order: bit-reproducible, no host sync) and grads are multiplied by
 * grad_mul * min(1, max_norm / (norm * grad_mul + 1e-6)); the last double of the workspace receives norm^2.
 * step_dev / lr_dev (device pointers, may be NULL): when given, the step count (bias corrections, computed on the device
 * in double) and the per-group learning rates are READ FROM DEVICE MEMORY instead of the host arrays, so a captured CUDA
 * graph of the training step stays valid from step to step. */
size_t tcvn_adamw_workspace_bytes(void);
int tcvn_adamw_fused(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const uint8_t* select,
                     int n_groups, const double* lr, const double* beta1, const double* beta2, const double* eps,
                     const double* weight_decay, const int64_t* step, float max_norm, float grad_mul, void* workspace,
                     size_t workspace_bytes, const int64_t* step_dev, const float* lr_dev, tcvn_stream_t stream);
/* Every seeded kernel (dropout masks, training pixel noise) launched from the calling host thread after this call adds
 * *device_ptr to the seed argument it was given (NULL: off).  With the step counter in device memory a replayed CUDA
 * graph draws fresh masks; forward and backward of one step read the same value. */
int tcvn_set_seed_offset(const uint64_t* device_ptr);

/* tcgen05 kernels of the bf16 training path on caller-provided row-major bf16 matrices (csrc/umma_train.cu).
 * Weight gradient with MN-major operands (the reduction runs over pixel rows):
 *   dw[item][k][n] += sum_m act(A[m + item_shift[item], item_col[item] + k]) * G[m, g_col0 + n],  k, n in [0,128)
 * dw is fp32 [n_items][128][128], accumulated (every persistent CTA stores one partial sum into the workspace, a
 * reduction kernel adds them to dw); rows k >= item_valid[item] receive zero; rows of A
 * outside [0, rows) read as zero; act = PReLU(scale*x + shift) from a_fold = [scale | shift | alpha] (each
 * a_fold_cols wide; columns >= a_fold_cols are forced to zero) or identity when a_fold is NULL.
 * Replaces autograd's weight gradient of nn.Conv2d (dense_net.py:21-38). */
size_t tcvn_t_umma_wgrad_workspace_bytes(int n_items);
int tcvn_t_umma_wgrad(const void* a_bf16, int64_t rows, int a_cols, int a_pitch, int n_items, const int32_t* item_col,
                      const int32_t* item_shift, const int32_t* item_valid, const float* a_fold, int a_fold_cols,
                      const void* g_bf16, int g_cols, int g_pitch, int g_col0, float* dw, void* workspace,
                      size_t workspace_bytes, tcvn_stream_t stream);
/* Input gradient of the 3x3 convolution as a 3-tap shifted GEMM:
 *   out[p][c] = sum_dy sum_k g2x[p + (1 - dy) * ring_wp][k] * wd[dy][c][k]   (ring rows of out are written as zero)
 * g2x bf16 [rows][128] (columns dx*32 + n hold G[p + 1 - dx][n]; 96.. zero), wd bf16 [3][128][128], out bf16 [rows][128] */
int tcvn_t_umma_conv2_dgrad(const void* g2x_bf16, const void* wd_bf16, int64_t rows, int ring_hp, int ring_wp,
                            void* out_bf16, tcvn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training walks: train-mode forward (batch-statistics BatchNorm with running-buffer update, dropout) and the
 * hand-written backward of the two halves of the network.  Replaces torch autograd over
 * layers/dense_net.py:8-167 and networks/neutrino_full_base_network.py:99-188 in the reference's training_step
 * (trainers/neutrino_full_base_trainer.py:162-192).  `arena` / `grad_arena` hold the fp32 tensors of the
 * sub-module in state_dict order (running buffers included; their gradient slots are never touched); parameter
 * gradients are ACCUMULATED.  The workspace carries the saved activations from forward to backward and must not
 * be reused in between.  Dropout masks derive from (seed, site, element): pass the same values to both calls. */
/* prec: TCVN_FP32 = CUDA-core parity path; TCVN_BF16 = bf16 activations, every dense-block / transition convolution
 * (forward, input gradient, weight gradient) on tcgen05, fp32 parameters / statistics / parameter gradients */
size_t tcvn_cnn_train_workspace_bytes(const tcvn_cnn_desc* d, tcvn_precision prec, int n_images);
int tcvn_cnn_train_forward(const tcvn_cnn_desc* d, tcvn_precision prec, float* arena, const float* pixels, int n_images,
                           float p_drop, float momentum, uint64_t seed, uint64_t site, float* embedding, void* workspace,
                           size_t workspace_bytes, tcvn_stream_t stream);
/* d_embedding (n_images, out_features) is overwritten */
int tcvn_cnn_train_backward(const tcvn_cnn_desc* d, tcvn_precision prec, const float* arena, float* grad_arena,
                            const float* pixels, int n_images, float p_drop, uint64_t seed, uint64_t site, float* d_embedding,
                            void* workspace, size_t workspace_bytes, tcvn_stream_t stream);
size_t tcvn_seq_train_workspace_bytes(const tcvn_seq_desc* d, int n_events, int max_prongs, int n_prongs);
/* prong_logits: (max_prongs * n_events, classes) in (slot, event) order */
int tcvn_seq_train_forward(const tcvn_seq_desc* d, const float* position, float* combined, const float* encoder,
                           const float* event_decoder, float* prong_decoder, const float* event_embedding,
                           const float* prong_embedding, const uint8_t* event_mask, const uint8_t* prong_mask, int n_events,
                           int max_prongs, int n_prongs, float p_drop, float momentum, uint64_t seed, float* event_logits,
                           float* prong_logits, void* workspace, size_t workspace_bytes, tcvn_stream_t stream);
/* d_event_logits / d_prong_logits are overwritten; d_event_embedding (B, pixel+feature), d_prong_embedding (T, pixel) */
int tcvn_seq_train_backward(const tcvn_seq_desc* d, const float* position, const float* combined, const float* encoder,
                            const float* event_decoder, const float* prong_decoder, float* g_position, float* g_combined,
                            float* g_encoder, float* g_event_decoder, float* g_prong_decoder, const uint8_t* prong_mask,
                            int n_events, int max_prongs, int n_prongs, float p_drop, uint64_t seed, float* d_event_logits,
                            float* d_prong_logits, float* d_event_embedding, float* d_prong_embedding, void* workspace,
                            size_t workspace_bytes, tcvn_stream_t stream);

/* SM budget of the persistent tcgen05 kernels launched from the calling host thread until the next call (0 = the whole
 * chip).  Lets two independent walks on two streams (the event CNN and the prong CNN, neutrino_full_base_network.py:
 * 99-104) run side by side on disjoint SMs instead of time-slicing. */
int tcvn_set_sm_limit(int n_sms);

/* ------------------------------------------------------------------------------------------
 * Loss and validation metrics on the device: one launch each, no host sync.
 * tcvn_loss_forward replaces `loss` + the masked_select / log_softmax / softmax / argmax chain of `training_step`
 * (trainers/neutrino_full_base_trainer.py:148-192): focal loss -log p_t (1 - p_t)^gamma (gamma 0: cross entropy),
 * mean over the event rows and over the prong slots with target >= 0, out8 = {event_scale * event_loss +
 * prong_scale * prong_loss, event_loss, prong_loss, event accuracy, prong accuracy, event rows, selected prong
 * rows, 0}.  d_event_logits (B, E) / d_prong_logits (B, L, P) (both or neither) receive d total / d logits.
 * prong_logits[b][l][:] starts at element b * prong_stride_event + l * prong_stride_slot (the network returns a
 * transposed view); targets are int64, a prong target < 0 marks a padded slot. */
int tcvn_loss_forward(const float* event_logits, const int64_t* event_targets, int n_events, int event_classes,
                      const float* prong_logits, const int64_t* prong_targets, int max_prongs, int prong_classes,
                      int64_t prong_stride_event, int64_t prong_stride_slot, float gamma, float event_scale,
                      float prong_scale, float* out8, float* d_event_logits, float* d_prong_logits, tcvn_stream_t stream);
/* d_* = g_* times the device scalar *upstream (the chain rule through the scalar loss), one launch for both */
int tcvn_loss_backward(const float* upstream, const float* g_event, int64_t n_event, const float* g_prong, int64_t n_prong,
                       float* d_event_logits, float* d_prong_logits, tcvn_stream_t stream);
/* validation_step (trainers/neutrino_full_base_trainer.py:194-209): softmax probabilities of the event rows and of
 * the prong slots with target >= 0 (other slots are written as zeros) and the accuracy state
 * counters4 += {event hits, event rows, prong hits, selected prong rows} (torchmetrics Accuracy(task="multiclass")). */
int tcvn_metrics_update(const float* event_logits, const int64_t* event_targets, int n_events, int event_classes,
                        const float* prong_logits, const int64_t* prong_targets, int max_prongs, int prong_classes,
                        int64_t prong_stride_event, int64_t prong_stride_slot, int64_t* counters4, float* event_prob,
                        float* prong_prob, tcvn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * --sdxl pixel-map CNN (BASELINE configs[3]; transformercvn/network/layers/sdxl_net.py:7-42, whose arithmetic is
 * diffusers.models.vae.Encoder - un-vendored, unpinned third-party code: parity unpinned).  fp32 path:
 * the convolutions are tcvn_t_gemm over ringed channels-last maps; these are the kernels the DenseNet does not have. */
/* NCHW fp32 pixels -> ringed channels-last [n][(H+2)(W+2)][C], ring rows zero */
int tcvn_sdxl_pixels_to_ring(const float* pixels_nchw, int n, int C, int H, int W, float* out_ring, tcvn_stream_t stream);
/* torch.nn.GroupNorm(groups, C, eps) per image over the interior pixels (+ SiLU when silu != 0) of a ringed map;
 * ring rows of out are zero.  sums_workspace: 2 * n * groups doubles (overwritten).  Replaces nn.GroupNorm + nn.SiLU
 * of diffusers' ResnetBlock2D / Attention / conv_norm_out. */
int tcvn_sdxl_groupnorm(const float* x_ring, int n, int C, int groups, int H, int W, const float* gamma, const float* beta,
                        float eps, int silu, float* out_ring, double* sums_workspace, tcvn_stream_t stream);
/* patches of diffusers' Downsample2D (F.pad(x, (0,1,0,1)) -> conv3x3 stride 2): out ringed at (H/2, W/2) with 9*C
 * channels ordered (dy, dx, c), so that the convolution is a GEMM with K = 9*C */
int tcvn_sdxl_patch_s2(const float* x_ring, int n, int C, int H, int W, float* out_ring, tcvn_stream_t stream);

/* bf16 / tcgen05 path of the same CNN (csrc/sdxl16.cu, csrc/umma.cu): ringed channels-last bf16 maps [n*(H+2)*(W+2)][C].
 *   tcvn_sdxl16_patch27    conv_in's 3x3 patches of the NCHW fp32 pixels as one 64-wide bf16 K chunk: column (dy*3+dx)*3 + c
 *   tcvn_sdxl16_groupnorm  GroupNorm(1 group) (+ SiLU) per image, bit-reproducible statistics; C % 8 == 0
 *   tcvn_sdxl16_patch_s2   patches [rows_out][9*C] of the stride-2 convolution behind F.pad(x, (0,1,0,1))
 *   tcvn_sdxl16_conv       the convolution itself on the tensor cores, as a shifted GEMM:
 *        out[m, n] = sum_t sum_c A[m + tap_off[t], c] W[n, t*a_cols + c] + sum_c X2[m, c] W[n, n_taps*a_cols + c] + bias[n]
 *      (a 3x3 tap is a constant row offset of the ringed layout; X2 = the ResNet block's residual input, with identity or
 *      1x1-shortcut weights in its K range, so the residual add costs no extra pass; ring rows of out are written as
 *      zeros).  a_cols, x2_cols: multiples of 64.  W bf16 [n_tiles*128][n_taps*a_cols + x2_cols] K-major, zero rows
 *      beyond the real output width; bias / ones: fp32 [n_tiles*128].
 *   tcvn_sdxl16_to_f32     bf16 -> fp32 (the 1x1-spatial tail runs the fp32 kernels above) */
int tcvn_sdxl16_patch27(const float* pixels_nchw, int n, int C, int H, int W, float divisor, void* out_bf16, tcvn_stream_t stream);
size_t tcvn_sdxl16_groupnorm_workspace_bytes(int n);
int tcvn_sdxl16_groupnorm(const void* x_bf16, int n, int C, int H, int W, const float* gamma, const float* beta, float eps, int silu,
                          void* out_bf16, void* workspace, size_t workspace_bytes, tcvn_stream_t stream);
int tcvn_sdxl16_patch_s2(const void* x_bf16, int n, int C, int H, int W, void* out_bf16, tcvn_stream_t stream);
/* The same stride-2 convolution without materialised patches: tcvn_sdxl16_s2d writes the space-to-depth matrix
 * [n * (H/2+2) * (W/2+2)][4*C] (C % 64 == 0), tcvn_sdxl16_conv_s2 reads it as nine row-shifted, column-grouped views. */
int tcvn_sdxl16_s2d(const void* x_bf16, int n, int C, int H, int W, void* out_bf16, tcvn_stream_t stream);
int tcvn_sdxl16_conv_s2(const void* s2d_bf16, int64_t rows, int C, const void* w_bf16, int n_tiles, const float* bias_padded,
                        const float* ones_padded, void* out_bf16, int out_cols, int ring_hp, int ring_wp, tcvn_stream_t stream);
/* The 64 -> 64 channel 3x3 convolutions of the full-resolution stages (72 % of the network's FLOPs) as ONE fused tcgen05
 * kernel over 2-D tiles (csrc/umma_conv2d.cu):  out = conv3x3(silu(GroupNorm(x))) + bias (+ residual).  A haloed 18 x 10 pixel
 * patch is loaded once per 16 x 8 output tile through a 4-D tensor map, GroupNorm + SiLU are applied in place in shared
 * memory, the nine taps are nine UMMA descriptors into that patch, and the epilogue takes the statistics of the NEXT
 * GroupNorm from the values it stores (per-tile slots, fixed-order sum).
 *   in_stat [n] float2 (mean, rstd) of x (tcvn_sdxl16_gn_stats, or a previous call's out_stat); gamma / beta [64]
 *   w_bf16  [9*64][64]: row t*64 + n, column k = weight[n][k][t/3][t%3];  residual may be NULL
 *   stat_parts (may be NULL): tcvn_sdxl16_conv2d_stat_bytes(n, H, W) of scratch; out_stat (may be NULL): [n] float2 of out */
size_t tcvn_sdxl16_conv2d_stat_bytes(int n, int H, int W);
int tcvn_sdxl16_conv2d_c64(const void* x_bf16, int n, int H, int W, const void* in_stat, const float* gamma, const float* beta,
                           const void* w_bf16, const float* bias, const void* residual_bf16, void* out_bf16, void* stat_parts,
                           void* out_stat, float eps, tcvn_stream_t stream);
int tcvn_sdxl16_gn_stats(const void* x_bf16, int n, int C, int H, int W, float eps, void* out_stat, void* workspace,
                         size_t workspace_bytes, tcvn_stream_t stream);
int tcvn_sdxl16_to_f32(const void* x_bf16, int64_t count, float* out, tcvn_stream_t stream);
int tcvn_sdxl16_conv(const void* a_bf16, int64_t rows, int a_cols, int n_taps, const int32_t* tap_off, const void* x2_bf16,
                     int x2_cols, const void* w_bf16, int n_tiles, const float* bias_padded, const float* ones_padded,
                     void* out_bf16, int out_cols, int ring_hp, int ring_wp, tcvn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TCVN_H_ */
