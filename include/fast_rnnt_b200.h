/*
 * fast_rnnt_b200 — C ABI of the B200 (sm_100a) pruned RNN-T loss hot path.
 *
 * Drop-in boundary for Samsung/tf-fast-rnnt: these entry points are what a
 * TensorFlow custom-op shim (tf.load_op_library), a ctypes binding or any other
 * FFI binds instead of the reference's native entry points
 *   MutualInformationCuda / MutualInformationBackwardCuda / CumminCuda
 *   (tf_fast_rnnt/csrc/mutual_information.h:134-168)
 * and instead of the TensorFlow-graph math of
 *   tf_fast_rnnt/python/tf_fast_rnnt/rnnt_loss.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with host_;
 *     tensors are dense row-major, last axis unit-stride;
 *   - float tensors are float32; index tensors int32;
 *   - `boundary` is [B][4] = {s_begin, t_begin, s_end, t_end} per utterance
 *     (tf_fast_rnnt/python/tf_fast_rnnt/__init__.py:98-106) and is mandatory
 *     (README.md:5);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     nothing synchronises the host, nothing is allocated or freed by the
 *     library: the caller supplies `workspace` of at least the size the matching
 *     *_workspace_bytes() query returns (256-byte aligned);
 *   - return value: FRN_OK (0) or a negative FRN_E* code; CUDA launch errors are
 *     returned as FRN_ECUDA with the cudaError_t available from
 *     frn_last_cuda_error().  (The reference returns 1 unconditionally,
 *     mutual_information_cuda.cu:810,873,1011.)
 *   - re-entrant: no global mutable state apart from the per-thread last error.
 */
#ifndef FAST_RNNT_B200_H_
#define FAST_RNNT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRN_VERSION 100 /* 1.0.0 */

enum frn_status {
  FRN_OK = 0,
  FRN_EINVAL = -1,     /* bad shape / argument */
  FRN_EWORKSPACE = -2, /* workspace too small or misaligned */
  FRN_ECUDA = -3,      /* CUDA runtime error, see frn_last_cuda_error() */
  FRN_EUNSUPPORTED = -4
};

/* rnnt_type of the reference API (rnnt_loss.py:110-122) */
enum frn_rnnt_type { FRN_REGULAR = 0, FRN_MODIFIED = 1, FRN_CONSTRAINED = 2 };

/* element type of the joiner logits handed to the pruned loss */
enum frn_dtype { FRN_F32 = 0, FRN_BF16 = 1, FRN_F16 = 2 /* am / lm only (frn_cast_to_f32, *_lp) */ };

/* reduction of the reference API (rnnt_loss.py:327-338) */
enum frn_reduction { FRN_NONE = 0, FRN_MEAN = 1, FRN_SUM = 2 };

int frn_version(void);
const char *frn_status_string(int status);
int frn_last_cuda_error(void);
/* Number of kernels the library has launched (or recorded into a CUDA graph under
 * stream capture) in this process so far.  Counted only by the -DFRN_DEBUG_HOOKS build
 * (libfast_rnnt_b200_dbg.so: tests, the benchmark's kernels-per-step figure); the
 * product library keeps no such state and returns 0. */
unsigned long long frn_kernel_launches(void);

/* ------------------------------------------------------------------------
 * A4. Lattice recursion, forward + occupation counts in one call.
 * Replaces op "FastRNNTLoss" (tf_fast_rnnt_op.cc:27-34,48-117) =
 * MutualInformationCuda + 2 memsets + H2D copy + MutualInformationBackwardCuda.
 *   px [B][S][T1] with T1 == T+1 (regular) or T1 == T (modified recursion)
 *   py [B][S+1][T]
 *   ans [B]; px_grad [B][S][T1] (shape of px — reference defect D2 fixed);
 *   py_grad [B][S+1][T].  If calc_gradients == 0 the grad pointers may be NULL.
 * ---------------------------------------------------------------------- */
size_t frn_mi_workspace_bytes(int B, int S, int T, int T1);
int frn_mi_fwd_bwd(const float *px, const float *py, const int32_t *boundary,
                   int B, int S, int T, int T1, int calc_gradients, float *ans,
                   float *px_grad, float *py_grad, void *workspace,
                   size_t workspace_bytes, void *stream);

/* A4 on the pruning band (the recursion rnnt_loss_pruned runs, rnnt_loss.py:1116-1119,
 * without materialising the dense lattice the reference builds at :968-1013):
 *   pxc, pyc [B][T][R]   band log-probs: entry i of frame t is lattice row ranges[b,t,0]+i
 *                        (what frn_pruned_logprobs scatters into the dense px/py)
 *   ans [B]; pxc_grad, pyc_grad [B][T][R] occupation counts of the band arcs (NULL when
 *   calc_gradients == 0).  R <= 8 and frn_band_mi_workspace_bytes() != 0, else
 *   FRN_EUNSUPPORTED (use frn_pruned_logprobs + frn_mi_fwd_bwd). */
size_t frn_band_mi_workspace_bytes(int B, int S, int T, int R);
int frn_band_mi_fwd_bwd(const float *pxc, const float *pyc, const int32_t *ranges,
                        const int32_t *boundary, int B, int S, int T, int R,
                        int rnnt_type, float delay_penalty, int calc_gradients,
                        float *ans, float *pxc_grad, float *pyc_grad, void *workspace,
                        size_t workspace_bytes, void *stream);

/* Replaces op "Cummin" (tf_fast_rnnt_op.cc:36-38,135-165; CumminCuda):
 * inclusive running minimum along the last axis of an int32 [rows][n] matrix. */
int frn_cummin(const int32_t *in, int32_t *out, int rows, int n, void *stream);

/* ------------------------------------------------------------------------
 * A1/A2. get_rnnt_logprobs / get_rnnt_logprobs_smoothed
 * (rnnt_loss.py:63-223, 1132-1367): px [B][S][T1], py [B][S+1][T] in the
 * reference's layout.  smoothed != 0 selects the smoothed variant with the two
 * scales.  T1 = T+1 for FRN_REGULAR, T otherwise.
 * ---------------------------------------------------------------------- */
size_t frn_simple_logprobs_workspace_bytes(int B, int S, int T, int C);
int frn_simple_logprobs(const float *lm, const float *am, const int32_t *symbols,
                        const int32_t *boundary, int B, int S, int T, int C,
                        int termination_symbol, int rnnt_type, int smoothed,
                        float lm_only_scale, float am_only_scale, float *px,
                        float *py, void *workspace, size_t workspace_bytes,
                        void *stream);

/* ------------------------------------------------------------------------
 * A1/A2 + A3 + A4 fused: rnnt_loss_simple / rnnt_loss_smoothed
 * (rnnt_loss.py:225-338, 1369-1494) with reduction "none":
 *   scores[b] = p[b, s_end, t_end]  (loss = -scores; reduce with frn_reduce)
 *   px_grad [B][S][T1], py_grad [B][S+1][T] occupation counts (may be NULL when
 *   calc_gradients == 0).  The log-probs never leave the workspace.
 * ---------------------------------------------------------------------- */
size_t frn_simple_loss_workspace_bytes(int B, int S, int T, int C);
int frn_simple_loss(const float *lm, const float *am, const int32_t *symbols,
                    const int32_t *boundary, int B, int S, int T, int C,
                    int termination_symbol, int rnnt_type, int smoothed,
                    float lm_only_scale, float am_only_scale,
                    float delay_penalty, int calc_gradients, float *scores,
                    float *px_grad, float *py_grad, void *workspace,
                    size_t workspace_bytes, void *stream);

/* frn_simple_loss plus the am half of do_rnnt_pruning (am_pruned [B][T][R][C] = am broadcast over R,
 * rnnt_loss.py:802-806: it does not depend on the prune ranges) on `side_stream` BESIDE the lattice recursion:
 * forked behind the normaliser (fork_event) and joined back into `stream` after the read-out (join_event).
 * The recursion occupies 2 B of the 148 SMs; the copy's persistent CTAs (at most max_ctas, <= 0: 20) take others.
 * side_stream, fork_event, join_event: a second stream and two cudaEvent_t the caller owns; capturable in a CUDA
 * graph.  Then pass am_pruned = NULL to frn_do_pruning_add_joiner.  Shapes the fused path does not take (long
 * lattices on the row-scan kernel, C % 4 != 0) run the two calls one after the other on `stream`. */
int frn_simple_loss_bcast(const float *lm, const float *am, const int32_t *symbols,
                          const int32_t *boundary, int B, int S, int T, int C,
                          int termination_symbol, int rnnt_type, int smoothed,
                          float lm_only_scale, float am_only_scale, float delay_penalty,
                          int calc_gradients, float *scores, float *px_grad, float *py_grad,
                          int R, float *am_pruned, int max_ctas, void *side_stream,
                          void *fork_event, void *join_event, void *workspace,
                          size_t workspace_bytes, void *stream);

/* A9. Gradient of sum_b scores_grad[b]*scores[b] w.r.t. am [B][T][C] and lm
 * [B][S+1][C] given the occupation counts returned by frn_simple_loss with the
 * same arguments (what TensorFlow autodiff + _RNNTLossGrad, __init__.py:154-162,
 * produce for rnnt_loss_simple).  scores_grad == NULL means all ones.
 * frn_smoothed_loss_bwd: the same for rnnt_loss_smoothed (rnnt_loss.py:1266-1365),
 * including the path through the batch-global unigram (rnnt_loss.py:1279-1280) into
 * every lm row; occupation counts from frn_simple_loss(smoothed = 1) with the same
 * scales.  Both use frn_simple_loss_bwd_workspace_bytes(). */
size_t frn_simple_loss_bwd_workspace_bytes(int B, int S, int T, int C);
int frn_simple_loss_bwd(const float *lm, const float *am, const int32_t *symbols,
                        const int32_t *boundary, const float *px_grad,
                        const float *py_grad, const float *scores_grad, int B,
                        int S, int T, int C, int termination_symbol,
                        int rnnt_type, float *am_grad, float *lm_grad,
                        void *workspace, size_t workspace_bytes, void *stream);
int frn_smoothed_loss_bwd(const float *lm, const float *am, const int32_t *symbols,
                          const int32_t *boundary, const float *px_grad,
                          const float *py_grad, const float *scores_grad, int B,
                          int S, int T, int C, int termination_symbol,
                          int rnnt_type, float lm_only_scale, float am_only_scale,
                          float *am_grad, float *lm_grad, void *workspace,
                          size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * Batch sharded by utterance over several GPUs.  Every kernel works per utterance, with one
 * exception: rnnt_loss_smoothed's unigram is the mean of softmax(lm rows) over the WHOLE batch
 * (rnnt_loss.py:1279-1280).  For results identical to the unsharded batch:
 *   1. frn_smoothed_unigram_sums: sums[c] = sum over this rank's B (S+1) lm rows of softmax(row)[c], c < C,
 *      and sums[C] = B (S+1)  (workspace: frn_simple_logprobs_workspace_bytes(B, S, 1, C));
 *   2. the caller all-reduces (sum) the C+1 floats over the ranks (frn_allreduce_sum or its own collective);
 *   3. the *_sharded entry points take the all-reduced sums (NULL = this rank's batch is the whole batch).
 * Backward: d loss / d unigram has to be summed over the ranks before it reaches the lm rows -
 * frn_smoothed_loss_bwd_sharded(phase 1) leaves this rank's share in du [C], the caller all-reduces du,
 * phase 2 (same arguments and workspace) completes lm_grad.  phase 0 = both on one rank.
 * (Ranks must pad lm to the same S for the padded rows to enter the mean as they do unsharded.)
 * ---------------------------------------------------------------------- */
int frn_smoothed_unigram_sums(const float *lm, int B, int S, int C, float *sums,
                              void *workspace, size_t workspace_bytes, void *stream);
int frn_simple_logprobs_sharded(const float *lm, const float *am, const int32_t *symbols,
                                const int32_t *boundary, int B, int S, int T, int C,
                                int termination_symbol, int rnnt_type, int smoothed,
                                float lm_only_scale, float am_only_scale,
                                const float *unigram_sums, float *px, float *py,
                                void *workspace, size_t workspace_bytes, void *stream);
int frn_simple_loss_sharded(const float *lm, const float *am, const int32_t *symbols,
                            const int32_t *boundary, int B, int S, int T, int C,
                            int termination_symbol, int rnnt_type, int smoothed,
                            float lm_only_scale, float am_only_scale,
                            const float *unigram_sums, float delay_penalty,
                            int calc_gradients, float *scores, float *px_grad,
                            float *py_grad, void *workspace, size_t workspace_bytes,
                            void *stream);
int frn_smoothed_loss_bwd_sharded(const float *lm, const float *am, const int32_t *symbols,
                                  const int32_t *boundary, const float *px_grad,
                                  const float *py_grad, const float *scores_grad, int B,
                                  int S, int T, int C, int termination_symbol,
                                  int rnnt_type, float lm_only_scale, float am_only_scale,
                                  const float *unigram_sums, float *du, int phase,
                                  float *am_grad, float *lm_grad, void *workspace,
                                  size_t workspace_bytes, void *stream);
/*  The same two entry points for am / lm of element type `am_lm_dtype` (FRN_F32, FRN_BF16,
 * FRN_F16), consumed as they are: the row-statistics kernel is the only kernel that reads them (it widens every
 * element in registers and leaves the contraction's operands and the few gathered values the epilogue needs), so
 * mixed-precision encoders / decoders save the widening pass and the float32 copies.  unigram_sums as in the
 * *_sharded forms (NULL: this batch).  bf16 / fp16 need the tensor-core path (C % 4 == 0, 16-byte aligned bases):
 * FRN_EUNSUPPORTED otherwise - widen with frn_cast_to_f32 then.  Everything else (px, py, scores, gradients) stays
 * float32; the backward entry points take float32 am / lm. */
int frn_simple_logprobs_lp(const void *lm, const void *am, int am_lm_dtype, const int32_t *symbols,
                           const int32_t *boundary, int B, int S, int T, int C,
                           int termination_symbol, int rnnt_type, int smoothed,
                           float lm_only_scale, float am_only_scale, const float *unigram_sums,
                           float *px, float *py, void *workspace, size_t workspace_bytes,
                           void *stream);
int frn_simple_loss_lp(const void *lm, const void *am, int am_lm_dtype, const int32_t *symbols,
                       const int32_t *boundary, int B, int S, int T, int C,
                       int termination_symbol, int rnnt_type, int smoothed,
                       float lm_only_scale, float am_only_scale, const float *unigram_sums,
                       float delay_penalty, int calc_gradients, float *scores, float *px_grad,
                       float *py_grad, void *workspace, size_t workspace_bytes, void *stream);
/* In-place sum all-reduce of n floats over an NCCL communicator the caller owns (ncclComm_t as void*), enqueued
 * on `stream` - the scalar of reduction='sum'|'mean' and the C+1 / C floats above.  libnccl.so.2 is bound lazily
 * with dlopen (no link-time dependency; the instance already loaded in the process is the one found);
 * FRN_EUNSUPPORTED if it cannot be bound, FRN_ECUDA if NCCL reports an error. */
int frn_allreduce_sum(float *buf, size_t n, void *nccl_comm, void *stream);

/* ------------------------------------------------------------------------
 * A5. get_rnnt_prune_ranges (rnnt_loss.py:647-761): ranges [B][T][R_out],
 * R_out = (s_range > S ? S+1 : s_range), see frn_prune_ranges_width().
 * ---------------------------------------------------------------------- */
int frn_prune_ranges_width(int S, int s_range);
size_t frn_prune_ranges_workspace_bytes(int B, int T);
int frn_prune_ranges(const float *px_grad, const float *py_grad,
                     const int32_t *boundary, int B, int S, int T, int T1,
                     int s_range, int32_t *ranges, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * A6. do_rnnt_pruning (rnnt_loss.py:763-812) and its gradient.
 *   am [B][T][C], lm [B][S+1][C], ranges [B][T][R]
 *   am_pruned, lm_pruned [B][T][R][C]
 * ---------------------------------------------------------------------- */
int frn_do_pruning(const float *am, const float *lm, const int32_t *ranges, int B,
                   int S, int T, int R, int C, float *am_pruned,
                   float *lm_pruned, void *stream);
/*  The same for am / lm of element type `am_lm_dtype` (FRN_F32, FRN_BF16, FRN_F16); am_pruned /
 * lm_pruned have that type too, like the reference's broadcast_to / gather (rnnt_loss.py:802-811), which keep the
 * dtype of their inputs.  2-byte types need an even C (FRN_EUNSUPPORTED otherwise). */
int frn_do_pruning_lp(const void *am, const void *lm, int am_lm_dtype, const int32_t *ranges,
                      int B, int S, int T, int R, int C, void *am_pruned, void *lm_pruned,
                      void *stream);
int frn_do_pruning_bwd(const float *am_pruned_grad, const float *lm_pruned_grad,
                       const int32_t *ranges, int B, int S, int T, int R, int C,
                       float *am_grad, float *lm_grad, void *stream);
/* The additive joiner of the reference's tests/README (logits = am_pruned +
 * lm_pruned, simple_rnnt_loss_test.py:120-125) as an element-wise kernel over n
 * floats; stands in for the user's joiner network in the benchmark. */
int frn_add_joiner(const float *am_pruned, const float *lm_pruned, float *logits, size_t n,
                   void *stream);
/* do_rnnt_pruning and the additive joiner in one pass: am_pruned, lm_pruned AND
 * logits = am_pruned + lm_pruned [B][T][R][C] are all written, the sum taken from the
 * registers that hold the two rows (saves re-reading 2 x B T R C floats; bit-identical
 * to frn_do_pruning followed by frn_add_joiner).  am_pruned may be NULL (written by
 * frn_broadcast_am_pruned). */
int frn_do_pruning_add_joiner(const float *am, const float *lm, const int32_t *ranges,
                              int B, int S, int T, int R, int C, float *am_pruned,
                              float *lm_pruned, float *logits, void *stream);
/* The am half of do_rnnt_pruning alone, for overlap: am_pruned[b,t,i,:] = am[b,t,:] does not depend on
 * the ranges (rnnt_loss.py:802-806 broadcasts am before it gathers lm), so it may run on a second stream
 * beside frn_simple_loss / frn_prune_ranges, whose kernels are dependency-chain bound and leave SMs idle.
 * A persistent grid of at most max_ctas (<= 0: 20) single-warp CTAs driven by the bulk-copy engine; sized
 * not to take SMs or issue slots from the kernels it runs beside.  Then pass am_pruned = NULL to
 * frn_do_pruning_add_joiner (lm_pruned and logits only).  C % 4 == 0, 16-byte aligned pointers, else
 * FRN_EUNSUPPORTED. */
int frn_broadcast_am_pruned(const float *am, int B, int T, int R, int C, float *am_pruned,
                            int max_ctas, void *stream);
/* (f2) fused additive joiner: logits[b,t,i,:] = am[b,t,:] + lm[b,ranges[b,t,i],:]
 * without materialising am_pruned / lm_pruned. out_dtype: frn_dtype. */
int frn_pruned_add_joiner(const float *am, const float *lm, const int32_t *ranges,
                          int B, int S, int T, int R, int C, int out_dtype,
                          void *logits, void *stream);

/* ------------------------------------------------------------------------
 * A7. get_rnnt_logprobs_pruned (rnnt_loss.py:853-1020): dense px [B][S][T1],
 * py [B][S+1][T] from joiner logits [B][T][R][C] (float32 or bf16).
 * ---------------------------------------------------------------------- */
size_t frn_pruned_logprobs_workspace_bytes(int B, int S, int T, int R);
int frn_pruned_logprobs(const void *logits, int logits_dtype,
                        const int32_t *symbols, const int32_t *ranges,
                        const int32_t *boundary, int B, int S, int T, int R,
                        int C, int termination_symbol, int rnnt_type, float *px,
                        float *py, void *workspace, size_t workspace_bytes,
                        void *stream);

/* Backward of frn_pruned_logprobs (what TensorFlow autodiff derives through
 * rnnt_loss.py:942-1018): logits_grad [B][T][R][C] (dtype of logits) from the cotangents
 * px_grad [B][S][T1], py_grad [B][S+1][T] of the dense log-probs.  Entries the forward
 * overwrote with -inf pass nothing back.  Workspace: frn_pruned_logprobs_workspace_bytes(). */
int frn_pruned_logprobs_bwd(const void *logits, int logits_dtype,
                            const int32_t *symbols, const int32_t *ranges,
                            const int32_t *boundary, const float *px_grad,
                            const float *py_grad, int B, int S, int T, int R, int C,
                            int termination_symbol, int rnnt_type, void *logits_grad,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * A7 + A3 + A8 fused: rnnt_loss_pruned (rnnt_loss.py:1022-1130), reduction
 * "none", on the band only.  scores [B].  If logits_grad != NULL it receives
 * d(sum_b scores_grad[b]*scores[b])/d logits, same dtype/shape as logits
 * (scores_grad == NULL means all ones) — the backward TensorFlow autodiff runs
 * through rnnt_loss.py:942-1018.
 * ---------------------------------------------------------------------- */
size_t frn_pruned_loss_workspace_bytes(int B, int S, int T, int R);
/* The workspace the path that will actually run needs: with a band of R <= 8 and no (or a moderate) delay
 * penalty the recursion runs on the band itself and needs a few MB instead of the dense lattice's planes
 * (40 bytes per cell); frn_pruned_loss accepts any workspace of at least this size.
 * frn_pruned_loss_workspace_bytes() is the upper bound over all delay penalties. */
size_t frn_pruned_loss_min_workspace_bytes(int B, int S, int T, int R, float delay_penalty);
int frn_pruned_loss(const void *logits, int logits_dtype, const int32_t *symbols,
                    const int32_t *ranges, const int32_t *boundary, int B, int S,
                    int T, int R, int C, int termination_symbol, int rnnt_type,
                    float delay_penalty, const float *scores_grad, float *scores,
                    void *logits_grad, void *workspace, size_t workspace_bytes,
                    void *stream);

/* (f1) rnnt_loss on the full joiner, logits [B][T][S+1][C]
 * (rnnt_loss.py:340-551): the pruned loss with the identity band. */
size_t frn_joint_loss_workspace_bytes(int B, int S, int T);
int frn_joint_loss(const void *logits, int logits_dtype, const int32_t *symbols,
                   const int32_t *boundary, int B, int S, int T, int C,
                   int termination_symbol, int rnnt_type, float delay_penalty,
                   const float *scores_grad, float *scores, void *logits_grad,
                   void *workspace, size_t workspace_bytes, void *stream);

/* A3. out[0] = -sum(scores) (FRN_SUM) or -mean (FRN_MEAN); FRN_NONE writes
 * out[b] = -scores[b].  `denominator` <= 0 means B (pass the global batch size
 * when the sum is completed by an all-reduce across ranks). */
int frn_reduce(const float *scores, int B, int reduction, float denominator,
               float *out, void *stream);
/*  bf16 / fp16 am, lm (or logits) -> the float32 the kernels take:
 * dst[i] = (float)src[i], n elements, one streaming pass. */
int frn_cast_to_f32(const void *src, int src_dtype, size_t n, float *dst, void *stream);
/* The same for two score vectors of one step (simple and pruned loss) in one launch. */
int frn_reduce_pair(const float *scores_a, const float *scores_b, int B, int reduction,
                    float denominator, float *out_a, float *out_b, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FAST_RNNT_B200_H_ */
