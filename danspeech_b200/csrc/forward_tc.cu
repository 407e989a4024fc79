// bf16 tensor-core forward path (placeholder until the tcgen05 kernels land).
#include "model_types.cuh"
namespace dsb {
int finalize_tc(dsb_model*, cudaStream_t) { return set_error(DSB_ERR_UNSUPPORTED, "bf16 path not built yet"); }
size_t forward_tc_workspace_bytes(const dsb_model*, int, int) { return 0; }
int forward_tc(dsb_model*, const float*, const int32_t*, int, int, float*, int32_t*, void*, cudaStream_t) {
  return set_error(DSB_ERR_UNSUPPORTED, "bf16 path not built yet");
}
}  // namespace dsb
