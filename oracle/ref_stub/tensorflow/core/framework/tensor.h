// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for the part of TensorFlow's tensor.h that
// /root/reference/tf_fast_rnnt/csrc/mutual_information{.h,_cuda.cu} use:
// tensorflow::TTypes<T,N>::{Tensor,ConstTensor,Vec,Matrix,ConstMatrix}, i.e.
// Eigen row-major TensorMaps exposing dimension(i), operator()(i[,j[,k]]),
// data() and NumDimensions.  The maps are passed BY VALUE to the reference's
// kernels (mutual_information_cuda.cu:173,489), so this stays a POD.
#pragma once
#include <cassert>
#include <cstdint>
#include <functional>
#include <limits>
#include <type_traits>
#include <cuda_runtime.h>

namespace tensorflow {
template <typename T, int N>
struct StubMap {
  static constexpr int NumDimensions = N;
  T *ptr;
  int dims[N];
  __host__ __device__ int dimension(int i) const { return dims[i]; }
  __host__ __device__ T *data() const { return ptr; }
  __host__ __device__ T &operator()(int i) const { return ptr[i]; }
  __host__ __device__ T &operator()(int i, int j) const { return ptr[(size_t)i * dims[1] + j]; }
  __host__ __device__ T &operator()(int i, int j, int k) const {
    return ptr[((size_t)i * dims[1] + j) * dims[2] + k];
  }
};
template <typename T, int NDIMS = 1>
struct TTypes {
  typedef StubMap<T, NDIMS> Tensor;
  typedef StubMap<const T, NDIMS> ConstTensor;
  typedef StubMap<T, 1> Vec;
  typedef StubMap<T, 2> Matrix;
  typedef StubMap<const T, 2> ConstMatrix;
};
}  // namespace tensorflow
