// K4b wgrad on tcgen05 (MN-major TF32 operands). Placeholder until the tensor-core version lands:
// reports NT_ERR_UNSUPPORTED so nt_layer_backward_wgrad uses the fp32 FFMA kernel in gemm_simt.cu.
#include "common.cuh"

namespace nt {

size_t tc_wgrad_workspace_bytes(int64_t, int64_t) { return 0; }

int tc_layer_wgrad(const float*, const float*, const float*, const int32_t*, const int32_t*, int64_t, int64_t, int, float, float, uint64_t, uint64_t,
                   float*, float*, void*, size_t, int, cudaStream_t) {
  return NT_ERR_UNSUPPORTED;
}

}  // namespace nt
