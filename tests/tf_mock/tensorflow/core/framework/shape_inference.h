// see op_kernel.h in this directory
#pragma once
#include "tensorflow/core/framework/op_kernel.h"
namespace tensorflow {
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
struct DimensionOrConstant {            // as in TensorFlow: a dimension handle or an int64 constant
  DimensionOrConstant(DimensionHandle) {}
  DimensionOrConstant(long long) {}
};
class InferenceContext {
 public:
  template <typename T> Status GetAttr(const char *, T *) { return Status(); }
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  ShapeHandle UnknownShapeOfRank(int) { return ShapeHandle(); }
  ShapeHandle Vector(DimensionOrConstant) { return ShapeHandle(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  Status WithRank(ShapeHandle, int, ShapeHandle *) { return Status(); }
};
}  // namespace shape_inference
}  // namespace tensorflow
