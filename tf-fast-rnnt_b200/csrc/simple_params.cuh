// Parameters shared by the SIMT and the tensor-core implementation of the
// simple / smoothed log-probs contraction + epilogue.
#pragma once
#include "common.cuh"

namespace frn {
// Two-term float16 operands of the tensor-core normaliser, written by the row-statistics kernel:
// exp(x - rowmax) * 2^15 = h + l * 2^-11, planes [rows][Cp] of halves (Cp = C rounded up to 8: 16-byte row pitch
// for TMA; columns >= C are never read - the tensor maps end at C and TMA fills the rest of a box with zeros).
struct SplitPlanes {
  unsigned short *amh = nullptr, *aml = nullptr, *lmh = nullptr, *lml = nullptr;
  int Cp = 0;                           // 0: not wanted
};
// am[b,t,blank] [B][T], lm[b,s,blank] and lm[b,s,symbols[b,s]] [B][S+1], gathered by the row-statistics kernel
struct RowGathers {
  float *am_term = nullptr, *lm_term = nullptr, *lm_sym = nullptr;
  float *am_sum = nullptr, *lm_sum = nullptr;    // sum_c exp(x - rowmax) per am / lm row: the normaliser's accuracy guard
  int term = 0;
};
struct SimpleParams {
  const float *lm, *am;                 // float32 inputs (SIMT kernel); the tensor-core kernel never reads them
  const int32_t *symbols, *boundary;
  const float *lmmax, *ammax;           // row maxima
  const float *lmsum, *amonly, *logu;   // smoothed only (may be null)
  const float *pxam_t = nullptr;        // [B][T][S] am[b,t,symbols[b,s]], written by the row-statistics kernel
  SplitPlanes split;                    // tensor-core path: the contraction's operands
  RowGathers gat;                       // tensor-core path: the other am / lm values of the epilogue
  const void *am_raw = nullptr, *lm_raw = nullptr;   // am / lm as given (element type raw_dtype, frn_dtype): read only
  int raw_dtype = 0;                                  // by the guard's exact recomputation of a cell
  float *px, *py;                       // reference layout
  int B, S, T, T1, C, term, rnnt_type, smoothed;
  float comb, lm_scale, am_scale;       // 1-lm-am; scales with the 1e-20 substitution (rnnt_loss.py:1342-1349)
  // Arc-plane output (frn_simple_loss): instead of px/py the tensor-core kernel writes every live arc as
  // (mantissa, exponent) straight into the diagonal-major plane the wavefront recursion streams
  // (DpWorkspace::XY, mi_dp.cu) and fills the dead remainder of the plane; px/py are not touched.
  float4 *XY = nullptr;
  int P = 0, Dn = 0, k = 0;
  float delay_penalty = 0.f;            // rnnt_loss.py:316-321, folded into the symbol arcs
};
int launch_simple_logprobs_tc(const SimpleParams &sp, cudaStream_t stream);
// true when the tensor-core kernel can take this problem (TMA needs C % 4 == 0 and 16-byte aligned bases)
bool simple_logprobs_tc_applicable(const void *lm, const void *am, int C);
}  // namespace frn
