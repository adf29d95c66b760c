// Host-side launchers implemented next to their kernels; called from api.cu.
#pragma once
#include "common.cuh"

namespace frn {
// mi_dp.cu
int launch_skew_dense(const float *px, const float *py, const int32_t *boundary, const DpGeom &g,
                      const DpWorkspace &w, float delay_penalty, cudaStream_t stream);
int launch_chain(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, bool both_directions,
                 cudaStream_t stream);
int launch_finalize_dense(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, float *ans,
                          float *px_grad, float *py_grad, cudaStream_t stream);
// mi_scan.cu
bool scan_dp_feasible(int S, int T);
bool scan_dp_supported(int S, int T);
size_t scan_dp_workspace_bytes(int B, int S, int T);
int launch_scan_dp(const float *px, const float *py, const int32_t *boundary, int B, int S, int T, int T1,
                   float delay_penalty, bool want_grad, void *workspace, float *ans, float *px_grad,
                   float *py_grad, cudaStream_t stream);
// prune.cu
int launch_cummin(const int32_t *in, int32_t *out, int rows, int n, cudaStream_t stream);
int launch_prune_ranges(const float *px_grad, const float *py_grad, const int32_t *boundary, int B, int S, int T,
                        int T1, int R, int32_t *ranges, int32_t *s_begin_ws, cudaStream_t stream);
int launch_do_pruning(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                      float *am_p, float *lm_p, cudaStream_t stream);
int launch_do_pruning_add(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                          float *am_p, float *lm_p, float *logits, cudaStream_t stream);
int launch_broadcast_am(const float *am, int B, int T, int R, int C, float *am_p, int max_ctas, cudaStream_t stream);
int launch_do_pruning_bwd(const float *am_p_grad, const float *lm_p_grad, const int32_t *ranges, int B, int S,
                          int T, int R, int C, float *am_grad, float *lm_grad, cudaStream_t stream);
int launch_pruned_add_joiner(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R,
                             int C, int out_dtype, void *logits, cudaStream_t stream);
}  // namespace frn

namespace frn {
// collective.cu
int launch_allreduce_sum(float *buf, size_t n, void *comm, cudaStream_t stream);
// logprobs_simple.cu
size_t simple_stats_bytes(int B, int S, int T, int C);
// arcs != nullptr: the arcs go straight into the recursion's diagonal-major plane instead of px/py
struct ArcPlaneOut { float4 *XY; int P, Dn, k; float delay_penalty; };
bool simple_arc_plane_supported(const void *lm, const void *am, int C, int rnnt_type);
int launch_simple_logprobs(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                           int B, int S, int T, int C, int term, int rnnt_type, int smoothed,
                           float lm_only_scale, float am_only_scale, float *px, float *py, void *stats_ws,
                           cudaStream_t stream, const ArcPlaneOut *arcs = nullptr,
                           const float *unigram_sums = nullptr);
// the same for am / lm of element type `dtype` (frn_dtype); bf16 / fp16 need the tensor-core path (else FRN_EUNSUPPORTED)
int launch_simple_logprobs_any(const void *lm, const void *am, int dtype, const int32_t *symbols,
                               const int32_t *boundary, int B, int S, int T, int C, int term, int rnnt_type,
                               int smoothed, float lm_only_scale, float am_only_scale, float *px, float *py,
                               void *stats_ws, cudaStream_t stream, const ArcPlaneOut *arcs = nullptr,
                               const float *unigram_sums = nullptr);
int launch_smoothing_stats(const float *lm, const float *am, int B, int S, int T, int C, void *stats_ws,
                           cudaStream_t stream, const float *unigram_sums = nullptr);
int launch_unigram_sums(const float *lm, int B, int S, int C, void *stats_ws, float *sums, cudaStream_t stream);
// simple_bwd.cu
size_t simple_bwd_workspace_bytes(int B, int S, int T, int C);
int launch_simple_bwd(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                      const float *px_grad, const float *py_grad, const float *scores_grad, int B, int S, int T, int C,
                      int term, int rnnt_type, int smoothed, float lm_only_scale, float am_only_scale, float *am_grad,
                      float *lm_grad, void *workspace, cudaStream_t stream, const float *unigram_sums = nullptr,
                      float *du_io = nullptr, int phase = 0);
// logprobs_pruned.cu
int launch_pruned_lse(const void *logits, int dtype, const int32_t *symbols, const int32_t *ranges, int B, int S,
                      int T, int R, int C, int term, float *pxc, float *pyc, float *lse, cudaStream_t stream);
int launch_skew_band(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary,
                     const DpGeom &g, const DpWorkspace &w, int R, int rnnt_type, float delay_penalty,
                     cudaStream_t stream);
int launch_finalize_band(const int32_t *ranges, const int32_t *boundary, const DpGeom &g, const DpWorkspace &w,
                         int R, int rnnt_type, float *gxc, float *gyc, float *scores, cudaStream_t stream);
int launch_pruned_logits_grad(const void *logits, int dtype, const int32_t *symbols, const int32_t *ranges,
                              const float *lse, const float *gxc, const float *gyc, const float *scores_grad,
                              int B, int S, int T, int R, int C, int term, void *dlogits, cudaStream_t stream);
int launch_dense_to_band(const float *dpx, const float *dpy, const int32_t *ranges, const int32_t *boundary, int B,
                         int S, int T, int T1, int R, int rnnt_type, float *gxc, float *gyc, cudaStream_t stream);
int launch_band_to_dense(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary,
                         int B, int S, int T, int T1, int R, int rnnt_type, float *px, float *py,
                         cudaStream_t stream);
// band_dp.cu
size_t band_dp_workspace_bytes(int B, int T);
bool band_dp_supported(int S, int T, int R);
int launch_band_dp(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary, int B, int S,
                   int T, int R, int rnnt_type, float delay_penalty, bool want_grad, void *workspace, float *gxc,
                   float *gyc, float *scores, cudaStream_t stream);
// misc.cu
int launch_cast_to_f32(const void *src, int dtype, size_t n, float *dst, cudaStream_t stream);
int launch_reduce(const float *scores, int B, int reduction, float denom, float *out, cudaStream_t stream);
int launch_reduce_pair(const float *a, const float *b, int B, int reduction, float denom, float *out_a, float *out_b,
                       cudaStream_t stream);
int launch_add(const float *a, const float *b, float *out, size_t n, cudaStream_t stream);
int launch_iota_ranges(int32_t *ranges, size_t n, int R, cudaStream_t stream);
}  // namespace frn
