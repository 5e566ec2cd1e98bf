#include "umma.h"
namespace tcvn {
int umma_dense_layer(const CnnPlan&, const BlockPlan&, const LayerPlan&, const char*, void*, void*, long long, cudaStream_t) {
  return fail(TCVN_ERR_UNSUPPORTED, "bf16 tcgen05 path not built yet");
}
int umma_transition(const CnnPlan&, const BlockPlan&, const BlockPlan&, const char*, const void*, void*, long long, cudaStream_t) {
  return fail(TCVN_ERR_UNSUPPORTED, "bf16 tcgen05 path not built yet");
}
}  // namespace tcvn
