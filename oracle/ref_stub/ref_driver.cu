// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's native functions
// (MutualInformationCuda / MutualInformationBackwardCuda / CumminCuda,
// tf_fast_rnnt/csrc/mutual_information.h:134-168), compiled from the sources
// where they lie under /root/reference.  ref_fast_rnnt_loss() drives them the
// way the reference's op does (tf_fast_rnnt/python/csrc/tf_fast_rnnt_op.cc:
// 66-113): forward, two memsets, ans_grad = 1 upload, backward, stream sync.
// Used as the GPU oracle for the lattice recursion and as "the reference's own
// GPU op" timing baseline on the B200.
#include <vector>
#include "tf_fast_rnnt/csrc/mutual_information.h"

namespace tf = tensorflow;

extern "C" {

int ref_fast_rnnt_loss(const float *px, const float *py, const int32_t *boundary, int B, int S, int T, int T1,
                       int calc_gradients, float *p, float *ans, float *p_grad, float *px_grad, float *py_grad,
                       float *ans_grad, int px_grad_T1, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  tf::TTypes<float, 3>::ConstTensor px_t{px, {B, S, T1}};
  tf::TTypes<float, 3>::ConstTensor py_t{py, {B, S + 1, T}};
  tf::TTypes<int32_t>::ConstMatrix bd_t{boundary, {B, 4}};
  tf::TTypes<float, 3>::Tensor p_t{p, {B, S + 1, T + 1}};
  tf::TTypes<float>::Vec ans_t{ans, {B}};
  int status = tf_fast_rnnt::MutualInformationCuda<float>(px_t, py_t, bd_t, p_t, ans_t, stream);
  if (calc_gradients) {
    // the op allocates px_grad as [B,S,T+1] whatever px is (op.cc:84); the
    // caller chooses px_grad_T1 (T+1 reproduces the op, T1 is the sane shape)
    tf::TTypes<float, 3>::Tensor pg_t{p_grad, {B, S + 1, T + 1}};
    tf::TTypes<float, 3>::Tensor gx_t{px_grad, {B, S, px_grad_T1}};
    tf::TTypes<float, 3>::Tensor gy_t{py_grad, {B, S + 1, T}};
    tf::TTypes<float>::Vec ag_t{ans_grad, {B}};
    cudaMemsetAsync(px_grad, 0, sizeof(float) * (size_t)B * S * px_grad_T1, stream);
    cudaMemsetAsync(py_grad, 0, sizeof(float) * (size_t)B * (S + 1) * T, stream);
    std::vector<float> ones(B, 1.0f);
    cudaMemcpyAsync(ans_grad, ones.data(), sizeof(float) * B, cudaMemcpyHostToDevice, stream);
    tf_fast_rnnt::MutualInformationBackwardCuda<float>(px_t, py_t, bd_t, p_t, pg_t, gx_t, gy_t, ag_t, true,
                                                       stream);
  }
  cudaStreamSynchronize(stream);
  return status;
}

int ref_cummin(const int32_t *in, int32_t *out, int rows, int n, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  tf::TTypes<int32_t>::ConstMatrix in_t{in, {rows, n}};
  tf::TTypes<int32_t>::Matrix out_t{out, {rows, n}};
  int status = tf_fast_rnnt::CumminCuda<int32_t>(in_t, out_t, stream);
  cudaStreamSynchronize(stream);
  return status;
}

}  // extern "C"
