// Beam CTC decoder (placeholder; implemented after the offline path is green on the GPU).
#include "model_types.cuh"
using namespace dsb;
extern "C" int dsb_beam_create(const char*, int, const char*, float, float, int, float, int, int, int, dsb_beam**) {
  return set_error(DSB_ERR_UNSUPPORTED, "beam decoder not built yet");
}
extern "C" void dsb_beam_destroy(dsb_beam*) {}
extern "C" size_t dsb_beam_workspace_bytes(const dsb_beam*, int, int) { return 0; }
extern "C" int dsb_beam_decode(dsb_beam*, const float*, const int32_t*, int, int, int, int32_t*, int32_t*, float*,
                               int32_t*, void*, size_t, void*) {
  return set_error(DSB_ERR_UNSUPPORTED, "beam decoder not built yet");
}
extern "C" int dsb_beam_lm_order(const dsb_beam*) { return 0; }
extern "C" int dsb_beam_lm_is_char_based(const dsb_beam*) { return 0; }
extern "C" int64_t dsb_beam_lm_num_ngrams(const dsb_beam*) { return 0; }
