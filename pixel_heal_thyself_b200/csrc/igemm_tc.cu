// tcgen05 implicit-GEMM (placeholder until the tensor-core kernels land).
#include "common.cuh"
namespace pht {
int conv_gemm_tc(const pht_conv_gemm_args*, cudaStream_t, bool* handled) { *handled = false; return PHT_OK; }
int wgrad_tc(const pht_wgrad_args*, cudaStream_t, bool* handled) { *handled = false; return PHT_OK; }
size_t wgrad_tc_workspace_bytes(const pht_wgrad_args*) { return 0; }
}  // namespace pht
