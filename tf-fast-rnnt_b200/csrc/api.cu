// extern "C" entry points of libfast_rnnt_b200.so (see include/fast_rnnt_b200.h).
// Argument validation, workspace carve-up and kernel sequencing only; all
// kernels live in the other translation units.
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "launchers.h"

// NVTX range around every entry point that enqueues work (tracing).  Compile-time switch: the
// debug-hooks library is built with -DFRN_NVTX (Makefile), the product library carries no tracing code.
#ifdef FRN_NVTX
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define FRN_RANGE() NvtxRange frn_nvtx_range__(__func__)
#else
#define FRN_RANGE() do { } while (0)
#endif

namespace frn {
static thread_local int g_last_cuda_error = 0;
#ifdef FRN_DEBUG_HOOKS
static std::atomic<unsigned long long> g_kernel_launches{0};
void count_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
#endif
int note_cuda_error(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return e == cudaSuccess ? FRN_OK : FRN_ECUDA;
}
int check_launch() { return note_cuda_error(cudaPeekAtLastError()); }

static inline bool aligned256(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 255u) == 0; }
static inline int type_t1(int T, int rnnt_type) { return rnnt_type == FRN_REGULAR ? T + 1 : T; }
// The band recursion keeps the R rows of a column as float64 mantissas in one frame.  A delay penalty p puts
// e^(p (T/2 - t)) on every symbol arc, so rows k symbols apart differ by up to e^(k p T / 2) inside one column
// or chunk image, and products of two such vectors (image x boundary state, alpha x beta) by twice that:
// beyond 400 bits over R - 1 rows (measured: exact to 1e-6 at 230 bits, wrong at 570) the dense-lattice
// kernels, which carry one frame per row, take over.
// Which of the two dense-lattice recursions runs a shape, and the workspace that one needs (sized by the path
// that will actually run: the wavefront's diagonal-major planes are 40 bytes per cell, the row scan's 16).
static inline bool dense_dp_uses_scan(const DpGeom &g) {
  return scan_dp_supported(g.S, g.T) || (g.P > kMaxRowsDp && scan_dp_feasible(g.S, g.T));
}
static inline size_t dense_dp_workspace_bytes(const DpGeom &g) {
  if (dense_dp_uses_scan(g)) return scan_dp_workspace_bytes(g.B, g.S, g.T);
  return g.P > kMaxRowsDp ? 0 : carve_dp(nullptr, g).bytes;     // 0: no dense recursion for this shape
}
static inline bool band_delay_ok(int T, int R, float delay_penalty) {
  return !(delay_penalty > 0.f) || (double)delay_penalty * T * (R > 1 ? R - 1 : 1) * 0.5 * 1.4427 < 400.0;
}
}  // namespace frn

using namespace frn;

#define FRN_REQUIRE(cond) \
  do {                    \
    if (!(cond)) return FRN_EINVAL; \
  } while (0)
#define FRN_TRY(expr)      \
  do {                     \
    int rc__ = (expr);     \
    if (rc__) return rc__; \
  } while (0)

extern "C" {

int frn_version(void) { return FRN_VERSION; }

const char *frn_status_string(int status) {
  switch (status) {
    case FRN_OK: return "ok";
    case FRN_EINVAL: return "invalid argument";
    case FRN_EWORKSPACE: return "workspace too small or misaligned";
    case FRN_ECUDA: return "CUDA error";
    case FRN_EUNSUPPORTED: return "unsupported size";
    default: return "unknown status";
  }
}

int frn_last_cuda_error(void) { return g_last_cuda_error; }

unsigned long long frn_kernel_launches(void) {
#ifdef FRN_DEBUG_HOOKS
  return g_kernel_launches.load(std::memory_order_relaxed);
#else
  return 0;     // not counted in the product build
#endif
}

// ------------------------------------------------------------------ A4
size_t frn_mi_workspace_bytes(int B, int S, int T, int T1) {
  if (B <= 0 || S < 0 || T < 0) return 0;
  return dense_dp_workspace_bytes(make_geom(B, S, T, T1));
}

int frn_mi_fwd_bwd(const float *px, const float *py, const int32_t *boundary, int B, int S, int T, int T1,
                   int calc_gradients, float *ans, float *px_grad, float *py_grad, void *workspace,
                   size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FRN_REQUIRE(B > 0 && S >= 0 && T >= 0 && (T1 == T || T1 == T + 1));
  FRN_REQUIRE(px && py && boundary && ans);
  FRN_REQUIRE(!calc_gradients || (px_grad && py_grad));
  DpGeom g = make_geom(B, S, T, T1);
  if (!workspace || !aligned256(workspace) || workspace_bytes < frn_mi_workspace_bytes(B, S, T, T1)) return FRN_EWORKSPACE;
  // the row scan where it is the faster kernel, and for lattices with more rows than the wavefront holds
  if (dense_dp_uses_scan(g))
    return launch_scan_dp(px, py, boundary, B, S, T, T1, 0.f, calc_gradients != 0, workspace, ans, px_grad, py_grad,
                          stream);
  if (g.P > kMaxRowsDp) return FRN_EUNSUPPORTED;
  DpWorkspace w = carve_dp(workspace, g);
  // FRN_MI_PHASE = 1 / 2 / 3 (debug-hooks build): only the skew / recursion / read-out kernel, on the planes a
  // previous full call left in the workspace - how the benchmark times the dependency-chain kernel alone
  const int phase = debug_env_int("FRN_MI_PHASE", 0);
  if (phase == 0 || phase == 1) FRN_TRY(launch_skew_dense(px, py, boundary, g, w, 0.f, stream));
  if (phase == 0 || phase == 2) FRN_TRY(launch_chain(boundary, g, w, calc_gradients != 0, stream));
  if (phase == 0 || phase == 3)
    FRN_TRY(launch_finalize_dense(boundary, g, w, ans, calc_gradients ? px_grad : nullptr,
                                  calc_gradients ? py_grad : nullptr, stream));
  return FRN_OK;
}

size_t frn_band_mi_workspace_bytes(int B, int S, int T, int R) {
  if (B <= 0 || S < 0 || T <= 0 || R <= 0 || !band_dp_supported(S, T, R)) return 0;
  return band_dp_workspace_bytes(B, T);
}

int frn_band_mi_fwd_bwd(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary, int B,
                        int S, int T, int R, int rnnt_type, float delay_penalty, int calc_gradients, float *ans,
                        float *pxc_grad, float *pyc_grad, void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && R >= 1 && R <= S + 1);
  FRN_REQUIRE(pxc && pyc && ranges && boundary && ans);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  FRN_REQUIRE(!calc_gradients || (pxc_grad && pyc_grad));
  if (!band_dp_supported(S, T, R) || !band_delay_ok(T, R, delay_penalty)) return FRN_EUNSUPPORTED;
  if (!workspace || !aligned256(workspace) || workspace_bytes < band_dp_workspace_bytes(B, T)) return FRN_EWORKSPACE;
  return launch_band_dp(pxc, pyc, ranges, boundary, B, S, T, R, rnnt_type, delay_penalty > 0.f ? delay_penalty : 0.f,
                        calc_gradients != 0, workspace, pxc_grad, pyc_grad, ans, static_cast<cudaStream_t>(stream));
}

int frn_cummin(const int32_t *in, int32_t *out, int rows, int n, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(rows >= 0 && n >= 0 && (rows == 0 || n == 0 || (in && out)));
  return launch_cummin(in, out, rows, n, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ A5
int frn_prune_ranges_width(int S, int s_range) { return s_range > S ? S + 1 : s_range; }

size_t frn_prune_ranges_workspace_bytes(int B, int T) {
  return round_up_sz((size_t)(B > 0 ? B : 0) * (T > 0 ? T : 0) * sizeof(int32_t), 256);
}

int frn_prune_ranges(const float *px_grad, const float *py_grad, const int32_t *boundary, int B, int S, int T,
                     int T1, int s_range, int32_t *ranges, void *workspace, size_t workspace_bytes,
                     void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && (T1 == T || T1 == T + 1) && s_range >= 1);
  FRN_REQUIRE(px_grad && py_grad && boundary && ranges);
  if (!workspace || workspace_bytes < frn_prune_ranges_workspace_bytes(B, T)) return FRN_EWORKSPACE;
  const int R = frn_prune_ranges_width(S, s_range);
  return launch_prune_ranges(px_grad, py_grad, boundary, B, S, T, T1, R, ranges,
                             static_cast<int32_t *>(workspace), static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ A6
int frn_do_pruning(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                   float *am_pruned, float *lm_pruned, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 0 && T > 0 && R > 0 && C > 0);
  FRN_REQUIRE(am_pruned || lm_pruned);
  FRN_REQUIRE((!am_pruned || am) && (!lm_pruned || (lm && ranges)));
  return launch_do_pruning(am, lm, ranges, B, S, T, R, C, am_pruned, lm_pruned, static_cast<cudaStream_t>(stream));
}

// am / lm of a 2-byte element type: the gather is a copy of whole rows, so pairs of elements travel as one 32-bit
// word through the same kernels (C must be even; the 128-bit path then needs C % 8 == 0)
int frn_do_pruning_lp(const void *am, const void *lm, int am_lm_dtype, const int32_t *ranges, int B, int S, int T, int R,
                      int C, void *am_pruned, void *lm_pruned, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(am_lm_dtype == FRN_F32 || am_lm_dtype == FRN_BF16 || am_lm_dtype == FRN_F16);
  if (am_lm_dtype == FRN_F32)
    return frn_do_pruning(static_cast<const float *>(am), static_cast<const float *>(lm), ranges, B, S, T, R, C,
                          static_cast<float *>(am_pruned), static_cast<float *>(lm_pruned), stream);
  FRN_REQUIRE(B > 0 && S >= 0 && T > 0 && R > 0 && C > 0);
  FRN_REQUIRE(am_pruned || lm_pruned);
  FRN_REQUIRE((!am_pruned || am) && (!lm_pruned || (lm && ranges)));
  if (C % 2 != 0) return FRN_EUNSUPPORTED;
  return launch_do_pruning(static_cast<const float *>(am), static_cast<const float *>(lm), ranges, B, S, T, R, C / 2,
                           static_cast<float *>(am_pruned), static_cast<float *>(lm_pruned),
                           static_cast<cudaStream_t>(stream));
}

int frn_do_pruning_add_joiner(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R,
                              int C, float *am_pruned, float *lm_pruned, float *logits, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 0 && T > 0 && R > 0 && C > 0);
  FRN_REQUIRE(am && lm && ranges && lm_pruned && logits);      // am_pruned == NULL: see frn_broadcast_am_pruned
  return launch_do_pruning_add(am, lm, ranges, B, S, T, R, C, am_pruned, lm_pruned, logits,
                               static_cast<cudaStream_t>(stream));
}

int frn_broadcast_am_pruned(const float *am, int B, int T, int R, int C, float *am_pruned, int max_ctas,
                            void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && T > 0 && R > 0 && C > 0 && am && am_pruned);
  return launch_broadcast_am(am, B, T, R, C, am_pruned, max_ctas, static_cast<cudaStream_t>(stream));
}

int frn_do_pruning_bwd(const float *am_pruned_grad, const float *lm_pruned_grad, const int32_t *ranges, int B,
                       int S, int T, int R, int C, float *am_grad, float *lm_grad, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 0 && T > 0 && R > 0 && C > 0 && ranges);
  FRN_REQUIRE((!am_grad || am_pruned_grad) && (!lm_grad || lm_pruned_grad));
  return launch_do_pruning_bwd(am_pruned_grad, lm_pruned_grad, ranges, B, S, T, R, C, am_grad, lm_grad,
                               static_cast<cudaStream_t>(stream));
}

int frn_pruned_add_joiner(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R,
                          int C, int out_dtype, void *logits, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 0 && T > 0 && R > 0 && C > 0 && am && lm && ranges && logits);
  return launch_pruned_add_joiner(am, lm, ranges, B, S, T, R, C, out_dtype, logits,
                                  static_cast<cudaStream_t>(stream));
}


// ------------------------------------------------------------------ A1 / A2
size_t frn_simple_logprobs_workspace_bytes(int B, int S, int T, int C) {
  if (B <= 0 || S < 0 || T <= 0 || C <= 0) return 0;
  return simple_stats_bytes(B, S, T, C);
}

int frn_allreduce_sum(float *buf, size_t n, void *nccl_comm, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(buf && n > 0 && nccl_comm);
  return launch_allreduce_sum(buf, n, nccl_comm, static_cast<cudaStream_t>(stream));
}

int frn_smoothed_unigram_sums(const float *lm, int B, int S, int C, float *sums, void *workspace,
                              size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 0 && C >= 1 && lm && sums);
  if (!workspace || !aligned256(workspace) || workspace_bytes < simple_stats_bytes(B, S, 1, C)) return FRN_EWORKSPACE;
  return launch_unigram_sums(lm, B, S, C, workspace, sums, static_cast<cudaStream_t>(stream));
}

int frn_simple_logprobs(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary, int B,
                        int S, int T, int C, int termination_symbol, int rnnt_type, int smoothed,
                        float lm_only_scale, float am_only_scale, float *px, float *py, void *workspace,
                        size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  return frn_simple_logprobs_sharded(lm, am, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                                     lm_only_scale, am_only_scale, nullptr, px, py, workspace, workspace_bytes, stream);
}

int frn_simple_logprobs_sharded(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                                int B, int S, int T, int C, int termination_symbol, int rnnt_type, int smoothed,
                                float lm_only_scale, float am_only_scale, const float *unigram_sums, float *px,
                                float *py, void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  return frn_simple_logprobs_lp(lm, am, FRN_F32, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                                lm_only_scale, am_only_scale, unigram_sums, px, py, workspace, workspace_bytes, stream);
}

int frn_simple_logprobs_lp(const void *lm, const void *am, int am_lm_dtype, const int32_t *symbols,
                           const int32_t *boundary, int B, int S, int T, int C, int termination_symbol, int rnnt_type,
                           int smoothed, float lm_only_scale, float am_only_scale, const float *unigram_sums, float *px,
                           float *py, void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(am_lm_dtype == FRN_F32 || am_lm_dtype == FRN_BF16 || am_lm_dtype == FRN_F16);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1);
  FRN_REQUIRE(lm && am && symbols && boundary && px && py);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < simple_stats_bytes(B, S, T, C)) return FRN_EWORKSPACE;
  return launch_simple_logprobs_any(lm, am, am_lm_dtype, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type,
                                    smoothed, lm_only_scale, am_only_scale, px, py, workspace,
                                    static_cast<cudaStream_t>(stream), nullptr, smoothed ? unigram_sums : nullptr);
}

// ------------------------------------------------------------------ A1/A2 + A3 + A4
namespace {
struct SimpleLossWs {
  float *px, *py;
  void *stats;
  void *dp;
  size_t bytes;
};
SimpleLossWs carve_simple_loss(void *base, int B, int S, int T, int T1, int C) {
  SimpleLossWs w;
  char *p = static_cast<char *>(base);
  w.px = reinterpret_cast<float *>(p); p += round_up_sz((size_t)B * S * T1 * sizeof(float), 256);
  w.py = reinterpret_cast<float *>(p); p += round_up_sz((size_t)B * (S + 1) * T * sizeof(float), 256);
  w.stats = p; p += simple_stats_bytes(B, S, T, C);
  w.dp = p; p += dense_dp_workspace_bytes(make_geom(B, S, T, T1));
  w.bytes = (size_t)(p - static_cast<char *>(base));
  return w;
}
}  // namespace

size_t frn_simple_loss_workspace_bytes(int B, int S, int T, int C) {
  if (B <= 0 || S < 0 || T <= 0 || C <= 0) return 0;
  return carve_simple_loss(nullptr, B, S, T, T + 1, C).bytes;  // T1 = T+1 is the larger case
}

int frn_simple_loss(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary, int B,
                    int S, int T, int C, int termination_symbol, int rnnt_type, int smoothed,
                    float lm_only_scale, float am_only_scale, float delay_penalty, int calc_gradients,
                    float *scores, float *px_grad, float *py_grad, void *workspace, size_t workspace_bytes,
                    void *stream_) {
  FRN_RANGE();
  return frn_simple_loss_sharded(lm, am, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                                 lm_only_scale, am_only_scale, nullptr, delay_penalty, calc_gradients, scores, px_grad,
                                 py_grad, workspace, workspace_bytes, stream_);
}

int frn_simple_loss_sharded(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary, int B,
                            int S, int T, int C, int termination_symbol, int rnnt_type, int smoothed,
                            float lm_only_scale, float am_only_scale, const float *unigram_sums, float delay_penalty,
                            int calc_gradients, float *scores, float *px_grad, float *py_grad, void *workspace,
                            size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  return frn_simple_loss_lp(lm, am, FRN_F32, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                            lm_only_scale, am_only_scale, unigram_sums, delay_penalty, calc_gradients, scores, px_grad,
                            py_grad, workspace, workspace_bytes, stream_);
}

int frn_simple_loss_lp(const void *lm, const void *am, int am_lm_dtype, const int32_t *symbols,
                       const int32_t *boundary, int B, int S, int T, int C, int termination_symbol, int rnnt_type,
                       int smoothed, float lm_only_scale, float am_only_scale, const float *unigram_sums,
                       float delay_penalty, int calc_gradients, float *scores, float *px_grad, float *py_grad,
                       void *workspace, size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  FRN_REQUIRE(am_lm_dtype == FRN_F32 || am_lm_dtype == FRN_BF16 || am_lm_dtype == FRN_F16);
  if (!smoothed) unigram_sums = nullptr;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1);
  FRN_REQUIRE(lm && am && symbols && boundary && scores);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  FRN_REQUIRE(!calc_gradients || (px_grad && py_grad));
  const int T1 = type_t1(T, rnnt_type);
  DpGeom g = make_geom(B, S, T, T1);
  const bool scan = dense_dp_uses_scan(g);
  if (!scan && g.P > kMaxRowsDp) return FRN_EUNSUPPORTED;
  if (!workspace || !aligned256(workspace) || workspace_bytes < carve_simple_loss(nullptr, B, S, T, T1, C).bytes)
    return FRN_EWORKSPACE;
  SimpleLossWs w = carve_simple_loss(workspace, B, S, T, T1, C);
  DpWorkspace dw = scan ? DpWorkspace{} : carve_dp(w.dp, g);
  const float dp = delay_penalty > 0.f ? delay_penalty : 0.f;
  if (!scan && simple_arc_plane_supported(lm, am, C, rnnt_type)) {
    // 4 launches: row statistics, normaliser (arcs straight into the recursion's plane), recursion, read-out
    const ArcPlaneOut arcs{dw.XY, g.P, g.Dn, g.k, dp};
    FRN_TRY(launch_simple_logprobs_any(lm, am, am_lm_dtype, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type,
                                       smoothed, lm_only_scale, am_only_scale, nullptr, nullptr, w.stats, stream, &arcs,
                                       unigram_sums));
    FRN_TRY(launch_chain(boundary, g, dw, calc_gradients != 0, stream));
    return launch_finalize_dense(boundary, g, dw, scores, calc_gradients ? px_grad : nullptr,
                                 calc_gradients ? py_grad : nullptr, stream);
  }
  FRN_TRY(launch_simple_logprobs_any(lm, am, am_lm_dtype, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type,
                                     smoothed, lm_only_scale, am_only_scale, w.px, w.py, w.stats, stream, nullptr,
                                     unigram_sums));
  if (scan)
    return launch_scan_dp(w.px, w.py, boundary, B, S, T, T1, dp, calc_gradients != 0, w.dp, scores, px_grad, py_grad,
                          stream);
  FRN_TRY(launch_skew_dense(w.px, w.py, boundary, g, dw, dp, stream));
  FRN_TRY(launch_chain(boundary, g, dw, calc_gradients != 0, stream));
  FRN_TRY(launch_finalize_dense(boundary, g, dw, scores, calc_gradients ? px_grad : nullptr,
                                calc_gradients ? py_grad : nullptr, stream));
  return FRN_OK;
}

// frn_simple_loss with the am half of do_rnnt_pruning (am_pruned[b,t,i,:] = am[b,t,:], independent of the prune
// ranges) running on `side_stream` BESIDE the lattice recursion: forked behind the normaliser - the row statistics
// and the normaliser want the whole GPU, the recursion occupies 2 B of its 148 SMs - and joined back into `stream`
// after the read-out.  The copy's persistent single-warp CTAs ask for enough shared memory not to share an SM with a
// recursion CTA.  fork_event / join_event: two caller-owned cudaEvent_t (no timing needed); everything stays
// stream-ordered and capturable in a CUDA graph.  Shapes the fused arc-plane path does not take run the two calls
// one after the other on `stream`.
int frn_simple_loss_bcast(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary, int B,
                          int S, int T, int C, int termination_symbol, int rnnt_type, int smoothed,
                          float lm_only_scale, float am_only_scale, float delay_penalty, int calc_gradients,
                          float *scores, float *px_grad, float *py_grad, int R, float *am_pruned, int max_ctas,
                          void *side_stream, void *fork_event, void *join_event, void *workspace,
                          size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_), side = static_cast<cudaStream_t>(side_stream);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1 && R >= 1);
  FRN_REQUIRE(lm && am && symbols && boundary && scores && am_pruned);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  FRN_REQUIRE(!calc_gradients || (px_grad && py_grad));
  const int T1 = type_t1(T, rnnt_type);
  DpGeom g = make_geom(B, S, T, T1);
  const bool fused = side && fork_event && join_event && side != stream && !dense_dp_uses_scan(g) &&
                     g.P <= kMaxRowsDp && simple_arc_plane_supported(lm, am, C, rnnt_type) && C % 4 == 0;
  if (!fused) {
    int rc = frn_simple_loss(lm, am, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                             lm_only_scale, am_only_scale, delay_penalty, calc_gradients, scores, px_grad, py_grad,
                             workspace, workspace_bytes, stream_);
    if (rc) return rc;
    return frn_broadcast_am_pruned(am, B, T, R, C, am_pruned, 296, stream_);     // alone on the GPU: all SMs
  }
  if (!workspace || !aligned256(workspace) || workspace_bytes < carve_simple_loss(nullptr, B, S, T, T1, C).bytes)
    return FRN_EWORKSPACE;
  SimpleLossWs w = carve_simple_loss(workspace, B, S, T, T1, C);
  DpWorkspace dw = carve_dp(w.dp, g);
  const float dp = delay_penalty > 0.f ? delay_penalty : 0.f;
  const ArcPlaneOut arcs{dw.XY, g.P, g.Dn, g.k, dp};
  FRN_TRY(launch_simple_logprobs(lm, am, symbols, boundary, B, S, T, C, termination_symbol, rnnt_type, smoothed,
                                 lm_only_scale, am_only_scale, nullptr, nullptr, w.stats, stream, &arcs, nullptr));
  cudaEvent_t fork = static_cast<cudaEvent_t>(fork_event), join = static_cast<cudaEvent_t>(join_event);
  cudaError_t e = cudaEventRecord(fork, stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fork, 0);
  if (e != cudaSuccess) return note_cuda_error(e);
  FRN_TRY(launch_broadcast_am(am, B, T, R, C, am_pruned, max_ctas, side));
  e = cudaEventRecord(join, side);
  if (e != cudaSuccess) return note_cuda_error(e);
  FRN_TRY(launch_chain(boundary, g, dw, calc_gradients != 0, stream));
  FRN_TRY(launch_finalize_dense(boundary, g, dw, scores, calc_gradients ? px_grad : nullptr,
                                calc_gradients ? py_grad : nullptr, stream));
  e = cudaStreamWaitEvent(stream, join, 0);
  return e == cudaSuccess ? FRN_OK : note_cuda_error(e);
}

size_t frn_simple_loss_bwd_workspace_bytes(int B, int S, int T, int C) {
  if (B <= 0 || S < 0 || T <= 0 || C <= 0) return 0;
  return simple_bwd_workspace_bytes(B, S, T, C);
}

int frn_simple_loss_bwd(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                        const float *px_grad, const float *py_grad, const float *scores_grad, int B, int S, int T,
                        int C, int termination_symbol, int rnnt_type, float *am_grad, float *lm_grad,
                        void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1);
  FRN_REQUIRE(lm && am && symbols && boundary && px_grad && py_grad && am_grad && lm_grad);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < simple_bwd_workspace_bytes(B, S, T, C))
    return FRN_EWORKSPACE;
  return launch_simple_bwd(lm, am, symbols, boundary, px_grad, py_grad, scores_grad, B, S, T, C, termination_symbol,
                           rnnt_type, 0, 0.f, 0.f, am_grad, lm_grad, workspace, static_cast<cudaStream_t>(stream));
}

int frn_smoothed_loss_bwd(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                          const float *px_grad, const float *py_grad, const float *scores_grad, int B, int S, int T,
                          int C, int termination_symbol, int rnnt_type, float lm_only_scale, float am_only_scale,
                          float *am_grad, float *lm_grad, void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  return frn_smoothed_loss_bwd_sharded(lm, am, symbols, boundary, px_grad, py_grad, scores_grad, B, S, T, C,
                                       termination_symbol, rnnt_type, lm_only_scale, am_only_scale, nullptr, nullptr, 0,
                                       am_grad, lm_grad, workspace, workspace_bytes, stream);
}

int frn_smoothed_loss_bwd_sharded(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                                  const float *px_grad, const float *py_grad, const float *scores_grad, int B, int S,
                                  int T, int C, int termination_symbol, int rnnt_type, float lm_only_scale,
                                  float am_only_scale, const float *unigram_sums, float *du, int phase,
                                  float *am_grad, float *lm_grad, void *workspace, size_t workspace_bytes,
                                  void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(phase >= 0 && phase <= 2 && (phase == 0 || (unigram_sums && du)));
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1);
  FRN_REQUIRE(lm && am && symbols && boundary && px_grad && py_grad && am_grad && lm_grad);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < simple_bwd_workspace_bytes(B, S, T, C))
    return FRN_EWORKSPACE;
  return launch_simple_bwd(lm, am, symbols, boundary, px_grad, py_grad, scores_grad, B, S, T, C, termination_symbol,
                           rnnt_type, 1, lm_only_scale, am_only_scale, am_grad, lm_grad, workspace,
                           static_cast<cudaStream_t>(stream), unigram_sums, du, phase);
}

// ------------------------------------------------------------------ A7 / A8
namespace {
struct PrunedWs {
  float *pxc, *pyc, *lse, *gxc, *gyc;
  int32_t *ranges;  // only for the full joiner
  void *dp;
  size_t bytes;
};
// recursion workspace of the pruned loss: the band recursion's states where it can run (R <= 8), the dense
// wavefront's planes where it may have to (wide bands, the full joiner, a delay penalty beyond band_delay_ok)
enum PrunedDp { kNoDp, kDpAnyPath, kDpBandOnly };
size_t pruned_dp_bytes(int B, int S, int T, int T1, int R, PrunedDp which) {
  const size_t band = band_dp_supported(S, T, R) ? band_dp_workspace_bytes(B, T) : 0;
  if (which == kDpBandOnly) return band;
  const DpGeom g = make_geom(B, S, T, T1);
  const size_t dense = g.P > kMaxRowsDp ? 0 : carve_dp(nullptr, g).bytes;
  return std::max(band, dense);
}
PrunedWs carve_pruned(void *base, int B, int S, int T, int T1, int R, PrunedDp with_dp, bool with_ranges) {
  PrunedWs w;
  char *p = static_cast<char *>(base);
  const size_t n = round_up_sz((size_t)B * T * R * sizeof(float), 256);
  w.pxc = reinterpret_cast<float *>(p); p += n;
  w.pyc = reinterpret_cast<float *>(p); p += n;
  w.lse = reinterpret_cast<float *>(p); p += n;
  w.gxc = reinterpret_cast<float *>(p); p += n;
  w.gyc = reinterpret_cast<float *>(p); p += n;
  w.ranges = reinterpret_cast<int32_t *>(p);
  if (with_ranges) p += n;
  w.dp = p;
  if (with_dp != kNoDp) p += pruned_dp_bytes(B, S, T, T1, R, with_dp);
  w.bytes = (size_t)(p - static_cast<char *>(base));
  return w;
}

// `dp_bytes`: what the caller's workspace holds behind w.dp
int pruned_loss_impl(const void *logits, int dtype, const int32_t *symbols, const int32_t *ranges,
                     const int32_t *boundary, int B, int S, int T, int R, int C, int term, int rnnt_type,
                     float delay_penalty, const float *scores_grad, float *scores, void *logits_grad,
                     const PrunedWs &w, size_t dp_bytes, cudaStream_t stream) {
  const int T1 = type_t1(T, rnnt_type);
  DpGeom g = make_geom(B, S, T, T1);
  FRN_TRY(launch_pruned_lse(logits, dtype, symbols, ranges, B, S, T, R, C, term, w.pxc, w.pyc, w.lse, stream));
  // narrow bands: transfer-matrix recursion on the band itself (band_dp.cu);
  // FRN_BAND_DENSE=1 (debug-hooks build) forces the dense-lattice wavefront for A/B runs and cross-checks
  const bool force_dense = debug_env_int("FRN_BAND_DENSE", 0) == 1;
  if (!force_dense && band_dp_supported(S, T, R) && band_delay_ok(T, R, delay_penalty) &&
      band_dp_workspace_bytes(B, T) <= dp_bytes) {
    const bool want = logits_grad != nullptr;
    FRN_TRY(launch_band_dp(w.pxc, w.pyc, ranges, boundary, B, S, T, R, rnnt_type, delay_penalty > 0.f ? delay_penalty : 0.f,
                           want, w.dp, w.gxc, w.gyc, scores, stream));
    if (want)
      FRN_TRY(launch_pruned_logits_grad(logits, dtype, symbols, ranges, w.lse, w.gxc, w.gyc, scores_grad, B, S, T, R, C,
                                        term, logits_grad, stream));
    return FRN_OK;
  }
  // dense-lattice wavefront on the planes the band is scattered into
  if (g.P > kMaxRowsDp) return FRN_EUNSUPPORTED;
  if (carve_dp(nullptr, g).bytes > dp_bytes) return FRN_EWORKSPACE;
  DpWorkspace dw = carve_dp(w.dp, g);
  FRN_TRY(launch_skew_band(w.pxc, w.pyc, ranges, boundary, g, dw, R, rnnt_type,
                           delay_penalty > 0.f ? delay_penalty : 0.f, stream));
  const bool want_grad = logits_grad != nullptr;
  FRN_TRY(launch_chain(boundary, g, dw, want_grad, stream));
  FRN_TRY(launch_finalize_band(ranges, boundary, g, dw, R, rnnt_type, want_grad ? w.gxc : nullptr,
                               want_grad ? w.gyc : nullptr, scores, stream));
  if (want_grad)
    FRN_TRY(launch_pruned_logits_grad(logits, dtype, symbols, ranges, w.lse, w.gxc, w.gyc, scores_grad, B, S, T,
                                      R, C, term, logits_grad, stream));
  return FRN_OK;
}
}  // namespace

size_t frn_pruned_logprobs_workspace_bytes(int B, int S, int T, int R) {
  if (B <= 0 || S < 0 || T <= 0 || R <= 0) return 0;
  return carve_pruned(nullptr, B, S, T, T + 1, R, kNoDp, false).bytes;
}

int frn_pruned_logprobs(const void *logits, int logits_dtype, const int32_t *symbols, const int32_t *ranges,
                        const int32_t *boundary, int B, int S, int T, int R, int C, int termination_symbol,
                        int rnnt_type, float *px, float *py, void *workspace, size_t workspace_bytes,
                        void *stream_) {
  FRN_RANGE();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && R >= 1 && C >= 1 && R <= S + 1);
  FRN_REQUIRE(logits && symbols && ranges && boundary && px && py);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < frn_pruned_logprobs_workspace_bytes(B, S, T, R))
    return FRN_EWORKSPACE;
  const int T1 = type_t1(T, rnnt_type);
  PrunedWs w = carve_pruned(workspace, B, S, T, T1, R, kNoDp, false);
  FRN_TRY(launch_pruned_lse(logits, logits_dtype, symbols, ranges, B, S, T, R, C, termination_symbol, w.pxc,
                            w.pyc, w.lse, stream));
  return launch_band_to_dense(w.pxc, w.pyc, ranges, boundary, B, S, T, T1, R, rnnt_type, px, py, stream);
}

int frn_pruned_logprobs_bwd(const void *logits, int logits_dtype, const int32_t *symbols, const int32_t *ranges,
                            const int32_t *boundary, const float *px_grad, const float *py_grad, int B, int S, int T,
                            int R, int C, int termination_symbol, int rnnt_type, void *logits_grad, void *workspace,
                            size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && R >= 1 && C >= 1 && R <= S + 1);
  FRN_REQUIRE(logits && symbols && ranges && boundary && px_grad && py_grad && logits_grad);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < frn_pruned_logprobs_workspace_bytes(B, S, T, R))
    return FRN_EWORKSPACE;
  const int T1 = type_t1(T, rnnt_type);
  PrunedWs w = carve_pruned(workspace, B, S, T, T1, R, kNoDp, false);
  FRN_TRY(launch_pruned_lse(logits, logits_dtype, symbols, ranges, B, S, T, R, C, termination_symbol, w.pxc,
                            w.pyc, w.lse, stream));
  FRN_TRY(launch_dense_to_band(px_grad, py_grad, ranges, boundary, B, S, T, T1, R, rnnt_type, w.gxc, w.gyc, stream));
  return launch_pruned_logits_grad(logits, logits_dtype, symbols, ranges, w.lse, w.gxc, w.gyc, nullptr, B, S, T, R, C,
                                   termination_symbol, logits_grad, stream);
}

size_t frn_pruned_loss_workspace_bytes(int B, int S, int T, int R) {
  if (B <= 0 || S < 0 || T <= 0 || R <= 0) return 0;
  return carve_pruned(nullptr, B, S, T, T + 1, R, kDpAnyPath, false).bytes;
}

size_t frn_pruned_loss_min_workspace_bytes(int B, int S, int T, int R, float delay_penalty) {
  if (B <= 0 || S < 0 || T <= 0 || R <= 0) return 0;
  const bool band = band_dp_supported(S, T, R) && band_delay_ok(T, R, delay_penalty) &&
                    debug_env_int("FRN_BAND_DENSE", 0) != 1;
  return carve_pruned(nullptr, B, S, T, T + 1, R, band ? kDpBandOnly : kDpAnyPath, false).bytes;
}

int frn_pruned_loss(const void *logits, int logits_dtype, const int32_t *symbols, const int32_t *ranges,
                    const int32_t *boundary, int B, int S, int T, int R, int C, int termination_symbol,
                    int rnnt_type, float delay_penalty, const float *scores_grad, float *scores,
                    void *logits_grad, void *workspace, size_t workspace_bytes, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && R >= 1 && C >= 1 && R <= S + 1);
  FRN_REQUIRE(logits && symbols && ranges && boundary && scores);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) ||
      workspace_bytes < frn_pruned_loss_min_workspace_bytes(B, S, T, R, delay_penalty))
    return FRN_EWORKSPACE;
  PrunedWs w = carve_pruned(workspace, B, S, T, type_t1(T, rnnt_type), R, kNoDp, false);
  return pruned_loss_impl(logits, logits_dtype, symbols, ranges, boundary, B, S, T, R, C, termination_symbol,
                          rnnt_type, delay_penalty, scores_grad, scores, logits_grad, w, workspace_bytes - w.bytes,
                          static_cast<cudaStream_t>(stream));
}

size_t frn_joint_loss_workspace_bytes(int B, int S, int T) {
  if (B <= 0 || S < 0 || T <= 0) return 0;
  return carve_pruned(nullptr, B, S, T, T + 1, S + 1, kDpAnyPath, true).bytes;
}

int frn_joint_loss(const void *logits, int logits_dtype, const int32_t *symbols, const int32_t *boundary, int B,
                   int S, int T, int C, int termination_symbol, int rnnt_type, float delay_penalty,
                   const float *scores_grad, float *scores, void *logits_grad, void *workspace,
                   size_t workspace_bytes, void *stream_) {
  FRN_RANGE();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FRN_REQUIRE(B > 0 && S >= 1 && T >= 1 && C >= 1);
  FRN_REQUIRE(logits && symbols && boundary && scores);
  FRN_REQUIRE(termination_symbol >= 0 && termination_symbol < C);
  FRN_REQUIRE(rnnt_type >= FRN_REGULAR && rnnt_type <= FRN_CONSTRAINED);
  if (!workspace || !aligned256(workspace) || workspace_bytes < frn_joint_loss_workspace_bytes(B, S, T))
    return FRN_EWORKSPACE;
  const int R = S + 1;
  PrunedWs w = carve_pruned(workspace, B, S, T, type_t1(T, rnnt_type), R, kNoDp, true);
  FRN_TRY(launch_iota_ranges(w.ranges, (size_t)B * T * R, R, stream));
  return pruned_loss_impl(logits, logits_dtype, symbols, w.ranges, boundary, B, S, T, R, C, termination_symbol,
                          rnnt_type, delay_penalty, scores_grad, scores, logits_grad, w, workspace_bytes - w.bytes,
                          stream);
}

int frn_add_joiner(const float *am_pruned, const float *lm_pruned, float *logits, size_t n, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(am_pruned && lm_pruned && logits);
  return launch_add(am_pruned, lm_pruned, logits, n, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ A3
int frn_reduce(const float *scores, int B, int reduction, float denominator, float *out, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && scores && out && reduction >= FRN_NONE && reduction <= FRN_SUM);
  return launch_reduce(scores, B, reduction, denominator > 0.f ? denominator : (float)B, out,
                       static_cast<cudaStream_t>(stream));
}

int frn_cast_to_f32(const void *src, int src_dtype, size_t n, float *dst, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(n == 0 || (src && dst));
  FRN_REQUIRE(src_dtype == FRN_BF16 || src_dtype == FRN_F16);
  return launch_cast_to_f32(src, src_dtype, n, dst, static_cast<cudaStream_t>(stream));
}

int frn_reduce_pair(const float *scores_a, const float *scores_b, int B, int reduction, float denominator,
                    float *out_a, float *out_b, void *stream) {
  FRN_RANGE();
  FRN_REQUIRE(B > 0 && scores_a && scores_b && out_a && out_b && reduction >= FRN_NONE && reduction <= FRN_SUM);
  return launch_reduce_pair(scores_a, scores_b, B, reduction, denominator > 0.f ? denominator : (float)B, out_a, out_b,
                            static_cast<cudaStream_t>(stream));
}

}  // extern "C"
