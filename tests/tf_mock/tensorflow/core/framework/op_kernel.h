// Minimal stand-in for the part of the TensorFlow C++ op API that tf_shim/tf_fast_rnnt_b200_ops.cc
// uses, so that the shim (which cannot be built in this image: TensorFlow is absent) is at least
// parsed and type-checked by g++ -fsyntax-only (tests/test_cabi_cpu.py).  Test infrastructure only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <type_traits>
#include <vector>

namespace Eigen {
struct GpuDevice {
  void *stream() const { return nullptr; }
};
struct half { uint16_t x; };
}  // namespace Eigen

namespace tensorflow {
using int32 = int32_t;
using int64 = long long;
using uint8 = uint8_t;
struct bfloat16 { uint16_t v; };
enum DataType { DT_FLOAT, DT_INT32, DT_UINT8, DT_BFLOAT16 };
constexpr const char *DEVICE_GPU = "GPU";

class Status {
 public:
  bool ok() const { return true; }
};
inline Status OkStatus() { return Status(); }
namespace errors {
template <typename... A> Status InvalidArgument(A &&...) { return Status(); }
template <typename... A> Status Internal(A &&...) { return Status(); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
 private:
  std::vector<int64> dims_;
};

template <typename T> struct Flat {
  T *data() const { return nullptr; }
};
template <typename T> struct Scalar {
  T operator()() const { return T(); }
};
class Tensor {
 public:
  int dims() const { return 0; }
  int64 dim_size(int) const { return 0; }
  const TensorShape &shape() const { return shape_; }
  int64 NumElements() const { return 0; }
  template <typename T> Flat<T> flat() const { return Flat<T>(); }
  template <typename T> Scalar<T> scalar() const { return Scalar<T>(); }
 private:
  TensorShape shape_;
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char *, T *) { return Status(); }
  void CtxFailure(const Status &) {}
};
class OpKernelContext {
 public:
  const Tensor &input(int) { return t_; }
  Status allocate_output(int, const TensorShape &, Tensor **out) { *out = &t_; return Status(); }
  Status allocate_temp(DataType, const TensorShape &, Tensor *) { return Status(); }
  template <typename D> const D &eigen_device() const { static D d; return d; }
  void CtxFailure(const Status &) {}
 private:
  Tensor t_;
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction *) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext *ctx) = 0;
};

#define OP_REQUIRES_OK(CTX, ...)                  \
  do {                                            \
    ::tensorflow::Status _s(__VA_ARGS__);         \
    if (!_s.ok()) { (CTX)->CtxFailure(_s); return; } \
  } while (0)
#define OP_REQUIRES(CTX, EXP, STATUS)             \
  do {                                            \
    if (!(EXP)) { (CTX)->CtxFailure(STATUS); return; } \
  } while (0)

#define TF_RETURN_IF_ERROR(...)                   \
  do {                                            \
    ::tensorflow::Status _s(__VA_ARGS__);         \
    if (!_s.ok()) return _s;                      \
  } while (0)

struct KernelDefBuilder {
  KernelDefBuilder &Device(const char *) { return *this; }
  KernelDefBuilder &HostMemory(const char *) { return *this; }
  template <typename T> KernelDefBuilder &TypeConstraint(const char *) { return *this; }
};
inline KernelDefBuilder Name(const char *) { return KernelDefBuilder(); }
#define TF_MOCK_CAT2(a, b) a##b
#define TF_MOCK_CAT(a, b) TF_MOCK_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(BUILDER, ...)                                                  \
  static ::tensorflow::OpKernel *TF_MOCK_CAT(tf_mock_make_, __LINE__)(::tensorflow::OpKernelConstruction *c) { \
    using namespace ::tensorflow;                                                              \
    (void)(BUILDER);                                                                           \
    return new __VA_ARGS__(c);                                                                 \
  }
}  // namespace tensorflow
