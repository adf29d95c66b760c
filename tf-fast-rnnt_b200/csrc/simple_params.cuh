// Parameters shared by the SIMT and the tensor-core implementation of the
// simple / smoothed log-probs contraction + epilogue.
#pragma once
#include "common.cuh"

namespace frn {
struct SimpleParams {
  const float *lm, *am;
  const int32_t *symbols, *boundary;
  const float *lmmax, *ammax;           // row maxima
  const float *lmsum, *amonly, *logu;   // smoothed only (may be null)
  float *px, *py;                       // reference layout
  int B, S, T, T1, C, term, rnnt_type, smoothed;
  float comb, lm_scale, am_scale;       // 1-lm-am; scales with the 1e-20 substitution (rnnt_loss.py:1342-1349)
};
int launch_simple_logprobs_tc(const SimpleParams &sp, cudaStream_t stream);
}  // namespace frn
