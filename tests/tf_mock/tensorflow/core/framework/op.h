// see op_kernel.h in this directory
#pragma once
#include "tensorflow/core/framework/shape_inference.h"
namespace tensorflow {
struct OpDefBuilderMock {
  OpDefBuilderMock &Input(const char *) { return *this; }
  OpDefBuilderMock &Output(const char *) { return *this; }
  OpDefBuilderMock &Attr(const char *) { return *this; }
  template <typename F> OpDefBuilderMock &SetShapeFn(F f) {
    Status (*fn)(shape_inference::InferenceContext *) = f;   // must be convertible like TF's OpShapeInferenceFn
    (void)fn;
    return *this;
  }
};
#define REGISTER_OP(NAME) static ::tensorflow::OpDefBuilderMock TF_MOCK_CAT(tf_mock_op_, __LINE__) = ::tensorflow::OpDefBuilderMock()
}  // namespace tensorflow
