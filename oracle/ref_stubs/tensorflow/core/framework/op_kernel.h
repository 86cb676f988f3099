// TEST INFRASTRUCTURE: a minimal stand-in for the TensorFlow op-kernel API, just enough to compile the
// reference's models/tf_ops/nn_distance/tf_nndistance.cpp and models/tf_ops/approxmatch/tf_approxmatch.cpp *where they
// lie* (TensorFlow is not installable here) and run their CPU kernels (NnDistanceOp, ApproxMatchOp, MatchCostOp ::Compute)
// as the oracles of the Chamfer-distance and EMD evaluators.
#pragma once
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <string>
#include <vector>

namespace tensorflow {

struct Status {
  bool ok_ = true;
  std::string msg;
  bool ok() const { return ok_; }
  static Status OK() { return Status(); }
};
namespace errors {
inline Status InvalidArgument(const char* m) { Status s; s.ok_ = false; s.msg = m; return s; }
}  // namespace errors

struct TensorShape {
  std::vector<int64_t> d;
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t> l) : d(l) {}
  int64_t dim_size(int i) const { return d[i]; }
  int dims() const { return (int)d.size(); }
  bool operator==(const TensorShape& o) const { return d == o.d; }
  int64_t num_elements() const { int64_t n = 1; for (auto v : d) n *= v; return n; }
};

template <class T>
struct Flat {
  T* p;
  T& operator()(int64_t i) const { return p[i]; }
};

struct Tensor {
  TensorShape shp;
  std::vector<unsigned char> buf;
  Tensor() {}
  Tensor(const TensorShape& s, size_t elem) : shp(s), buf((size_t)s.num_elements() * elem) {}
  int dims() const { return shp.dims(); }
  const TensorShape& shape() const { return shp; }
  template <class T> Flat<T> flat() { return Flat<T>{reinterpret_cast<T*>(buf.data())}; }
  template <class T> Flat<const T> flat() const { return Flat<const T>{reinterpret_cast<const T*>(buf.data())}; }
};

template <class T> struct DataTypeToEnum { static constexpr int value = 0; };

struct OpKernelConstruction {};
struct OpKernelContext {
  std::vector<Tensor> inputs;
  std::vector<Tensor*> outputs;
  Status status;
  const Tensor& input(int i) { return inputs[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if ((int)outputs.size() <= i) outputs.resize(i + 1, nullptr);
    outputs[i] = new Tensor(s, 4);     // float32 / int32 outputs only
    *t = outputs[i];
    return Status::OK();
  }
  Status allocate_temp(int, const TensorShape& s, Tensor* t) { *t = Tensor(s, 4); return Status::OK(); }
  ~OpKernelContext() { for (auto* t : outputs) delete t; }
};
struct OpKernel {
  explicit OpKernel(OpKernelConstruction*) {}
  virtual void Compute(OpKernelContext*) = 0;
  virtual ~OpKernel() {}
};

#define OP_REQUIRES(ctx, cond, st) do { if (!(cond)) { (ctx)->status = (st); return; } } while (0)
#define OP_REQUIRES_OK(ctx, expr) do { ::tensorflow::Status _s = (expr); if (!_s.ok()) { (ctx)->status = _s; return; } } while (0)

struct KernelDefBuilderStub { KernelDefBuilderStub& Device(const char*) { return *this; } };
inline KernelDefBuilderStub Name(const char*) { return KernelDefBuilderStub(); }
#define DEVICE_CPU "CPU"
#define DEVICE_GPU "GPU"
#define REGISTER_KERNEL_BUILDER(builder, cls) static const int _reg_kernel_##cls = ((void)(builder), 0)

}  // namespace tensorflow
