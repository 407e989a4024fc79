// Streaming forward (placeholder; implemented after the offline path is green on the GPU).
#include "model_types.cuh"
using namespace dsb;
extern "C" int dsb_stream_state_create(dsb_model*, int, int, dsb_stream_state**) {
  return set_error(DSB_ERR_UNSUPPORTED, "streaming not built yet");
}
extern "C" void dsb_stream_state_destroy(dsb_stream_state*) {}
extern "C" int dsb_stream_max_out_frames(const dsb_stream_state*, int) { return 0; }
extern "C" int dsb_streaming_forward(dsb_model*, dsb_stream_state*, const float*, int, int, int, float*, int32_t*,
                                     void*) {
  return set_error(DSB_ERR_UNSUPPORTED, "streaming not built yet");
}
