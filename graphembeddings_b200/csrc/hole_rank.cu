// placeholder, replaced below
#include "hole_common.cuh"
void hole_rank_ws_free(hole_ctx*) {}
extern "C" int hole_rank(hole_ctx*, const float*, int64_t, int64_t, const int32_t*, int64_t, int, int,
                         const int64_t*, const int32_t*, float*, int, int32_t*, int32_t*, void*) {
  return hole_set_error(HOLE_ERR_UNSUPPORTED, "hole_rank not built yet");
}
