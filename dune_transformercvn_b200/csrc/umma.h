// bf16 tcgen05 path: launch wrappers used by the CNN walker.
#pragma once
#include "kernels.h"

namespace tcvn {
int umma_dense_layer(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* packed, void* blk, void* mid,
                     long long rows, cudaStream_t st);
int umma_dense_layer_part(const CnnPlan& P, const BlockPlan& B, const LayerPlan& L, const char* packed, void* blk,
                          void* mid, long long rows, int which, cudaStream_t st);
int umma_transition(const CnnPlan& P, const BlockPlan& B, const BlockPlan& Nx, const char* packed, const void* pool,
                    void* next_blk, long long rows, cudaStream_t st);
}  // namespace tcvn
