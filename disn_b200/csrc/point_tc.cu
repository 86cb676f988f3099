// tcgen05 (bf16x3) fused point kernel -- placeholder until the tensor-core path lands.
#include "common.cuh"
namespace disn {
int tc_pack_weights(disn_ctx*) { return 0; }
int launch_point_tc(disn_ctx*, const PointJob&) {
  set_error("DISN_PREC_BF16X3 path is not built in this revision");
  return -4;
}
}  // namespace disn
