// Small utility kernels: loss reduction (A3) and identity band for the full joiner.
#include "common.cuh"
#include <cuda_fp16.h>

namespace frn {

// out = -scores (none) | -sum | -sum/denominator; one block, deterministic order.
// Block 1 (frn_reduce_pair) does the same for a second vector.
__global__ void __launch_bounds__(256) reduce_kernel(const float *scores, int B, int reduction, float denom,
                                                     float *out, const float *scores2, float *out2) {
  if (blockIdx.x == 1) { scores = scores2; out = out2; }
  if (reduction == FRN_NONE) {
    for (int i = threadIdx.x; i < B; i += blockDim.x) out[i] = -scores[i];
    return;
  }
  __shared__ float part[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) acc += scores[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += part[j];
    out[0] = (reduction == FRN_MEAN) ? -(s / denom) : -s;
  }
}

__global__ void iota_ranges_kernel(int32_t *ranges, size_t n, int R) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ranges[i] = (int32_t)(i % (size_t)R);
}

// out = a + b, the additive joiner of the reference's tests and README
// (simple_rnnt_loss_test.py:120-125: logits = pruned_am + pruned_lm)
__global__ void __launch_bounds__(256) add_kernel(const float *a, const float *b, float *out, size_t n) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 x = ld_stream_f4(reinterpret_cast<const float4 *>(a + i4));
    const float4 y = ld_stream_f4(reinterpret_cast<const float4 *>(b + i4));
    st_stream_f4(reinterpret_cast<float4 *>(out + i4), make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
  } else {
    for (size_t i = i4; i < n; ++i) out[i] = a[i] + b[i];
  }
}

// bf16 / fp16 -> float32, 8 elements (16 bytes in, 32 bytes out) per thread
template <typename T>
__global__ void __launch_bounds__(256) cast_to_f32_kernel(const T *src, float *dst, size_t n) {
  const size_t i8 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i8 + 7 < n) {
    uint4 raw;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(src + i8));
    const T *e = reinterpret_cast<const T *>(&raw);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = static_cast<float>(e[j]);
    st_stream_f4(reinterpret_cast<float4 *>(dst + i8), make_float4(v[0], v[1], v[2], v[3]));
    st_stream_f4(reinterpret_cast<float4 *>(dst + i8 + 4), make_float4(v[4], v[5], v[6], v[7]));
  } else {
    for (size_t i = i8; i < n; ++i) dst[i] = static_cast<float>(src[i]);
  }
}

int launch_cast_to_f32(const void *src, int dtype, size_t n, float *dst, cudaStream_t stream) {
  if (n == 0) return FRN_OK;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) return FRN_EINVAL;
  const unsigned grid = (unsigned)((n / 8 + 256) / 256);
  if (dtype == FRN_BF16)
    count_launch(), cast_to_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), dst, n);
  else if (dtype == FRN_F16)
    count_launch(), cast_to_f32_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(src), dst, n);
  else return FRN_EINVAL;
  return check_launch();
}

int launch_add(const float *a, const float *b, float *out, size_t n, cudaStream_t stream) {
  if (n == 0) return FRN_OK;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u)
    return FRN_EINVAL;
  count_launch(), add_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, stream>>>(a, b, out, n);
  return check_launch();
}

int launch_reduce(const float *scores, int B, int reduction, float denom, float *out, cudaStream_t stream) {
  count_launch(), reduce_kernel<<<1, 256, 0, stream>>>(scores, B, reduction, denom, out, nullptr, nullptr);
  return check_launch();
}

int launch_reduce_pair(const float *a, const float *b, int B, int reduction, float denom, float *out_a, float *out_b,
                       cudaStream_t stream) {
  count_launch(), reduce_kernel<<<2, 256, 0, stream>>>(a, B, reduction, denom, out_a, b, out_b);
  return check_launch();
}

int launch_iota_ranges(int32_t *ranges, size_t n, int R, cudaStream_t stream) {
  count_launch(), iota_ranges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ranges, n, R);
  return check_launch();
}

}  // namespace frn
