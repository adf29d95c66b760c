// Parameters shared by the kernels of the am / lm gradient (A9): simple_bwd.cu (weights, scatter terms,
// smoothing terms, SIMT contraction) and simple_bwd_tc.cu (tcgen05 contraction).
#pragma once
#include "common.cuh"

namespace frn {
struct BwdParams {
  const float *lm, *am;
  const int32_t *symbols, *boundary;
  const float *gpx, *gpy;       // occupation counts, reference layout ([B,S,T1], [B,S+1,T])
  const float *py;              // forward py [B,S+1,T] (recomputed), gives Z
  const float *lmmax, *ammax;   // row maxima
  const float *scores_grad;     // [B] or null (ones)
  float *W;                     // [B][S1p][Tp] workspace, zero outside [S+1][T] (S1p, Tp: multiples of 128)
  int S1p, Tp;
  float *am_grad, *lm_grad;
  int B, S, T, T1, C, term, rnnt_type;
  // smoothed loss only
  int smoothed;
  float comb, lm_scale, am_scale;           // 1 - lm - am; scales with the 1e-20 substitution
  const float *lmsum, *amonly, *unigram;    // forward statistics: sum_c exp(lm - lmmax); log D + ammax; u[c]
  const float *usums;                       // sharded batch: all-reduced unigram sums [C+1] (count last), or null
  float *Gt, *Sx, *Sy, *du, *partial;       // [B][T], [B][S+1], [B][S+1], [C], [chunks][C]
};
// tcgen05 contraction: am_grad / lm_grad = -g comb probs(x) * (W-weighted sums); FRN_EUNSUPPORTED when the
// shape needs the SIMT kernel (C % 4 != 0, misaligned bases)
int launch_bwd_contract_tc(const BwdParams &p, cudaStream_t stream);
}  // namespace frn
