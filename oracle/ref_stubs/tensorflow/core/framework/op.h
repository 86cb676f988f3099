// TEST INFRASTRUCTURE: stand-in for tensorflow/core/framework/op.h (see op_kernel.h in this directory).
#pragma once
namespace tensorflow {
struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
};
}  // namespace tensorflow
#define DISN_STUB_CAT2(a, b) a##b
#define DISN_STUB_CAT(a, b) DISN_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub DISN_STUB_CAT(_reg_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()
