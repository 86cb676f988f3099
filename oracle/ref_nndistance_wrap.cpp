// TEST INFRASTRUCTURE (oracle/_ref): compiles the REFERENCE's own Chamfer nearest-neighbour CPU op from the source
// where it lies (/root/reference/models/tf_ops/nn_distance/tf_nndistance.cpp, included below, never copied) against
// the stub TF headers in oracle/ref_stubs, and exposes NnDistanceOp::Compute through a C entry point.
#include <cstring>

#include REF_NNDISTANCE_CPP   // -DREF_NNDISTANCE_CPP="\"/root/reference/.../tf_nndistance.cpp\""

// the GPU launchers the reference file declares live in its .cu (not built here)
void NmDistanceKernelLauncher(int, int, const float*, int, const float*, float*, int*, float*, int*) {}
void NmDistanceGradKernelLauncher(int, int, const float*, int, const float*, const float*, const int*, const float*,
                                  const int*, float*, float*) {}

extern "C" int ref_nn_distance(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                               float* dist2, int* idx2, char* err, int errlen) {
  using namespace tensorflow;
  OpKernelConstruction c;
  NnDistanceOp op(&c);
  OpKernelContext ctx;
  ctx.inputs.emplace_back(TensorShape{b, n, 3}, 4);
  ctx.inputs.emplace_back(TensorShape{b, m, 3}, 4);
  std::memcpy(ctx.inputs[0].buf.data(), xyz1, (size_t)b * n * 3 * 4);
  std::memcpy(ctx.inputs[1].buf.data(), xyz2, (size_t)b * m * 3 * 4);
  op.Compute(&ctx);
  if (!ctx.status.ok()) {
    std::strncpy(err, ctx.status.msg.c_str(), errlen - 1);
    err[errlen - 1] = 0;
    return 1;
  }
  std::memcpy(dist1, ctx.outputs[0]->buf.data(), (size_t)b * n * 4);
  std::memcpy(idx1, ctx.outputs[1]->buf.data(), (size_t)b * n * 4);
  std::memcpy(dist2, ctx.outputs[2]->buf.data(), (size_t)b * m * 4);
  std::memcpy(idx2, ctx.outputs[3]->buf.data(), (size_t)b * m * 4);
  return 0;
}
