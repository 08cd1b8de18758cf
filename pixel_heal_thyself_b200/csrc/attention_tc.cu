// tcgen05 block-local attention (placeholder until the tensor-core kernels land).
#include "common.cuh"
namespace pht {
int attn_fwd_tc(const pht_attn_args*, cudaStream_t, bool* handled) { *handled = false; return PHT_OK; }
int attn_bwd_tc(const pht_attn_bwd_args*, cudaStream_t, bool* handled) { *handled = false; return PHT_OK; }
}  // namespace pht
