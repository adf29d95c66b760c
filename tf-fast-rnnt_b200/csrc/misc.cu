// Small utility kernels: loss reduction (A3) and identity band for the full joiner.
#include "common.cuh"

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
