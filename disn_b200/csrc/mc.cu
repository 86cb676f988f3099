// CUDA marching cubes -- placeholder until the post-pass lands.
#include "common.cuh"
namespace disn {
int marching_cubes(disn_ctx*, const float*, int, const double*, float, float*, int64_t*, int32_t*, int64_t*, bool) {
  set_error("marching cubes is not built in this revision");
  return -4;
}
}  // namespace disn
