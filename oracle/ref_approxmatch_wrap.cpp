// TEST INFRASTRUCTURE (oracle/_ref): compiles the REFERENCE's own approximate-EMD CPU ops from the source where it lies
// (/root/reference/models/tf_ops/approxmatch/tf_approxmatch.cpp, included below, never copied) against the stub TF headers
// in oracle/ref_stubs, and exposes ApproxMatchOp::Compute / MatchCostOp::Compute through C entry points.
#include <cstring>

#include REF_APPROXMATCH_CPP   // -DREF_APPROXMATCH_CPP="\"/root/reference/.../tf_approxmatch.cpp\""

// the GPU launchers the reference file declares live in its .cu (not built here)
void approxmatchLauncher(int, int, int, const float*, const float*, float*, float*) {}
void matchcostLauncher(int, int, int, const float*, const float*, const float*, float*) {}
void matchcostgradLauncher(int, int, int, const float*, const float*, const float*, float*, float*) {}

namespace {
int fail(const tensorflow::OpKernelContext& ctx, char* err, int errlen) {
  std::strncpy(err, ctx.status.msg.c_str(), errlen - 1);
  err[errlen - 1] = 0;
  return 1;
}
}  // namespace

// match: b * n * m floats, element (k, l) of batch item i at i*n*m + k*m + l (the CPU op's indexing)
extern "C" int ref_approx_match(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, char* err,
                                int errlen) {
  using namespace tensorflow;
  OpKernelConstruction c;
  ApproxMatchOp op(&c);
  OpKernelContext ctx;
  ctx.inputs.emplace_back(TensorShape{b, n, 3}, 4);
  ctx.inputs.emplace_back(TensorShape{b, m, 3}, 4);
  std::memcpy(ctx.inputs[0].buf.data(), xyz1, (size_t)b * n * 3 * 4);
  std::memcpy(ctx.inputs[1].buf.data(), xyz2, (size_t)b * m * 3 * 4);
  op.Compute(&ctx);
  if (!ctx.status.ok()) return fail(ctx, err, errlen);
  std::memcpy(match, ctx.outputs[0]->buf.data(), (size_t)b * n * m * 4);
  return 0;
}

extern "C" int ref_match_cost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* cost,
                              char* err, int errlen) {
  using namespace tensorflow;
  OpKernelConstruction c;
  MatchCostOp op(&c);
  OpKernelContext ctx;
  ctx.inputs.emplace_back(TensorShape{b, n, 3}, 4);
  ctx.inputs.emplace_back(TensorShape{b, m, 3}, 4);
  ctx.inputs.emplace_back(TensorShape{b, m, n}, 4);
  std::memcpy(ctx.inputs[0].buf.data(), xyz1, (size_t)b * n * 3 * 4);
  std::memcpy(ctx.inputs[1].buf.data(), xyz2, (size_t)b * m * 3 * 4);
  std::memcpy(ctx.inputs[2].buf.data(), match, (size_t)b * n * m * 4);
  op.Compute(&ctx);
  if (!ctx.status.ok()) return fail(ctx, err, errlen);
  std::memcpy(cost, ctx.outputs[0]->buf.data(), (size_t)b * 4);
  return 0;
}
