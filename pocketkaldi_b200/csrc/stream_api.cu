// Streaming entry points (carried state across chunks). Filled in after the batch
// path; the symbols exist so that the C ABI in include/pkb200.h is complete.

#include "nnet.cuh"

struct pkb_stream {
  pkb::Ctx *c = nullptr;
};

extern "C" {

int pkb_stream_create(pkb_ctx_t *, pkb_am_t *, int, int, const float *, float, pkb_stream_t **) {
  pkb::set_error("pkb_stream_create: streaming is not implemented yet");
  return PKB_ERR_UNSUPPORTED;
}
void pkb_stream_destroy(pkb_stream_t *st) { delete st; }
int pkb_stream_max_frames(const pkb_stream_t *) { return 0; }
int pkb_stream_push_i16(pkb_stream_t *, const int16_t *, float *, int32_t *) {
  pkb::set_error("pkb_stream_push_i16: streaming is not implemented yet");
  return PKB_ERR_UNSUPPORTED;
}
int pkb_stream_flush(pkb_stream_t *, float *, int32_t *) {
  pkb::set_error("pkb_stream_flush: streaming is not implemented yet");
  return PKB_ERR_UNSUPPORTED;
}

}  // extern "C"
