// Prune ranges (A5), cummin (op "Cummin"), pruning gather (A6) and its gradient.
//
// Replaces get_rnnt_prune_ranges / _adjust_pruning_lower_bound /
// _monotonic_lower_bound (rnnt_loss.py:553-761: ~15 TF kernels + two calls of
// the reference's scan kernel, mutual_information_cuda.cu:895-1012) and
// do_rnnt_pruning (rnnt_loss.py:763-812).  Integer results are bit-exact by
// construction: the float part walks s sequentially per (b,t) column, i.e. the
// summation order of a sequential cumsum, with first-index
// argmax; the int32 fix-ups along t are order independent.
#include <algorithm>

#include "common.cuh"
#include "launchers.h"

namespace frn {

// ---------------------------------------------------------------------------
// inclusive running minimum along the last axis, one warp per row
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_min_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = min(v, u);
  }
  return v;
}

__global__ void cummin_kernel(const int32_t *in, int32_t *out, int rows, int n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int32_t *src = in + (size_t)row * n;
  int32_t *dst = out + (size_t)row * n;
  int carry = INT32_MAX;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int v = (i < n) ? src[i] : INT32_MAX;
    v = min(warp_incl_min_scan(v, lane), carry);
    if (i < n) dst[i] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}

// ---------------------------------------------------------------------------
// A5 in two launches:
//  (1) prune_argmax_kernel, thread per (b,t) column, 128 columns per block:
//      s_begin[t] = argmax_k ( sum_{k<=s<k+R} py_grad[s,t] - px_grad[k-1,t] )
//      (rnnt_loss.py:722-748), coalesced along t, the window walked sequentially
//      in s (= the sequential-cumsum float order the oracle defines) in chunks of
//      16 with all loads of a chunk issued first;
//  (2) prune_fixup_kernel, one block per utterance: the monotonic fix-ups of
//      _adjust_pruning_lower_bound (:623-641) as two block-wide reverse running
//      minima on the int32 row held in shared memory, then
//      ranges[t,i] = s_begin[t] + i (:758-759).
// ---------------------------------------------------------------------------
constexpr int kPruneThreads = 512;

// reverse inclusive running minimum of x[0..T) in shared memory, whole block
__device__ void block_rev_cummin(int32_t *x, int T, int32_t *warp_carry) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = kPruneThreads / 32;
  int carry = INT32_MAX;  // running minimum of everything to the right of the current tile
  for (int hi = T - 1; hi >= 0; hi -= kPruneThreads) {
    const int i = hi - tid;  // thread 0 holds the largest t of the tile
    int v = (i >= 0) ? x[i] : INT32_MAX;
    v = warp_incl_min_scan(v, lane);
    if (lane == 31) warp_carry[w] = v;
    __syncthreads();
    int pre = carry;
    for (int j = 0; j < w; ++j) pre = min(pre, warp_carry[j]);
    v = min(v, pre);
    if (i >= 0) x[i] = v;
    int tile_min = warp_carry[0];
    for (int j = 1; j < NW; ++j) tile_min = min(tile_min, warp_carry[j]);
    carry = min(carry, tile_min);
    __syncthreads();
  }
}

constexpr int kArgmaxThreads = 128;   // maximum columns (t) per block; blockDim.x = columns actually used
#ifndef FRN_ARGMAX_TILE_KB
#define FRN_ARGMAX_TILE_KB 112        // 2 x 112 KB still fit one SM
#endif
constexpr size_t kArgmaxTileBytes = (size_t)FRN_ARGMAX_TILE_KB * 1024;

// The [S+1] x 128 tile of py_grad and the [S] x 128 tile of px_grad are brought into shared memory
// with 4-byte cp.async (fire and forget: every load of the tile is in flight at once, one memory
// round trip), then each thread walks its column sequentially out of shared memory.
__global__ void __launch_bounds__(kArgmaxThreads) prune_argmax_kernel(const float *px_grad, const float *py_grad,
                                                                      const int32_t *boundary, int S, int T, int T1,
                                                                      int R, int32_t *s_begin) {
  extern __shared__ float tile[];           // py: [S+1][cols], then px: [S][cols]
  const int b = blockIdx.y;
  const int tx = threadIdx.x, cols = blockDim.x;
  const int t = blockIdx.x * cols + tx;
  const int s_end = boundary[4 * b + 2], t_end = boundary[4 * b + 3];
  const int S1 = S + 1;
  const int nk = S1 - R + 1;
  float *spy = tile, *spx = tile + (size_t)S1 * cols;
  const bool live = t < T && t < t_end - 1;   // this column needs the arg-max (others are padding frames)
  if (live) {
    const float *py = py_grad + (size_t)b * S1 * T + t;
    const float *px = px_grad + (size_t)b * S * T1 + t;
    const uint32_t dpy = smem_u32(spy + tx), dpx = smem_u32(spx + tx);
    for (int s = 0; s < S1; ++s)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dpy + (uint32_t)(s * cols * 4)),
                   "l"(py + (size_t)s * T) : "memory");
    for (int s = 0; s < S; ++s)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dpx + (uint32_t)(s * cols * 4)),
                   "l"(px + (size_t)s * T1) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");   // each thread reads back only what it copied itself
  if (t >= T) return;
  int best_k = 0;
  if (live) {
    const float *cy = spy + tx, *cx = spx + tx;
    float cs_lo = 0.f, cs_hi = 0.f;
    for (int s = 0; s < R; ++s) cs_hi = cs_hi + cy[s * cols];  // cs[R], sequential
    float best = 0.f;
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
      float fin = cs_hi - cs_lo;
      if (k > 0) fin = fin - cx[(k - 1) * cols];
      if (k == 0 || fin > best) { best = fin; best_k = k; }
      cs_lo = cs_lo + cy[k * cols];
      if (k + R < S1) cs_hi = cs_hi + cy[(k + R) * cols];
    }
  } else {
    best_k = max(s_end - R + 1, 0);  // padding frames, rnnt_loss.py:744-748
  }
  s_begin[(size_t)b * T + t] = best_k;
}

__global__ void __launch_bounds__(kPruneThreads) prune_fixup_kernel(const int32_t *s_begin, int T, int R, int r_fix,
                                                                    int32_t *ranges) {
  extern __shared__ int32_t sb[];  // [T] s_begin, then [16] warp carries
  int32_t *warp_carry = sb + T;
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += kPruneThreads) sb[t] = s_begin[(size_t)b * T + t];
  __syncthreads();
  block_rev_cummin(sb, T, warp_carry);
  for (int t = threadIdx.x; t < T; t += kPruneThreads) sb[t] = -(sb[t] - (r_fix - 1) * t);
  __syncthreads();
  block_rev_cummin(sb, T, warp_carry);
  for (int t = threadIdx.x; t < T; t += kPruneThreads) sb[t] = -(max(sb[t], 0) - (r_fix - 1) * t);
  __syncthreads();
  int32_t *out = ranges + (size_t)b * T * R;
  for (int i = threadIdx.x; i < T * R; i += kPruneThreads) {
    const int t = i / R;
    out[i] = sb[t] + (i - t * R);
  }
}

// ---------------------------------------------------------------------------
// A6 forward: am_pruned[b,t,i,:] = am[b,t,:], lm_pruned[b,t,i,:] = lm[b,ranges[b,t,i],:]
// One warp per (b,t); 128-bit accesses when C % 4 == 0 (rows are then 16-byte
// aligned).  Out-of-range indices produce zeros (tf.gather on GPU).
// ---------------------------------------------------------------------------
// Vector path: one CTA of 128 threads per (b,t); thread <-> one float4 column
// (C <= 512 in one pass), all R source rows loaded before the 2R streaming stores.
// WITH_SUM: also logits[b,t,i,:] = am_pruned + lm_pruned (the additive joiner of the reference's tests) from
// the registers that hold both rows, instead of a second pass that re-reads the two pruned tensors.
template <int RMAX, bool WITH_AM, bool WITH_LM, bool WITH_SUM = false>
__global__ void __launch_bounds__(128) do_pruning_vec_kernel(const float *am, const float *lm, const int32_t *ranges,
                                                             int T, int S1, int R, int C4, float *am_p, float *lm_p,
                                                             float *sum_p = nullptr) {
  const int bt = blockIdx.x;
  const int b = bt / T;
  const int32_t *rg = ranges + (size_t)bt * R;
  const float4 *am_row = reinterpret_cast<const float4 *>(am) + (size_t)bt * C4;
  float4 *am_out = reinterpret_cast<float4 *>(am_p) + (size_t)bt * R * C4;
  float4 *lm_out = reinterpret_cast<float4 *>(lm_p) + (size_t)bt * R * C4;
  float4 *sum_out = reinterpret_cast<float4 *>(sum_p) + (size_t)bt * R * C4;
  const float4 *lm_b = reinterpret_cast<const float4 *>(lm) + (size_t)b * S1 * C4;
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WITH_AM || WITH_SUM) a = __ldg(am_row + c);
    float4 l[RMAX];
    if (WITH_LM) {
#pragma unroll
      for (int i = 0; i < RMAX; ++i) {
        if (i < R) {
          const int s = rg[i];
          l[i] = (s >= 0 && s < S1) ? __ldg(lm_b + (size_t)s * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) {
        if (WITH_AM) st_stream_f4(am_out + (size_t)i * C4 + c, a);
        if (WITH_LM) st_stream_f4(lm_out + (size_t)i * C4 + c, l[i]);
        if (WITH_SUM)
          st_stream_f4(sum_out + (size_t)i * C4 + c, make_float4(a.x + l[i].x, a.y + l[i].y, a.z + l[i].z, a.w + l[i].w));
      }
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) do_pruning_kernel(const float *am, const float *lm, const int32_t *ranges,
                                                         int BT, int T, int S1, int R, int C, float *am_p,
                                                         float *lm_p) {
  const int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bt >= BT) return;
  const int b = bt / T;
  const int32_t *rg = ranges + (size_t)bt * R;
  const float *am_row = am + (size_t)bt * C;
  float *am_out = am_p + (size_t)bt * R * C;
  float *lm_out = lm_p + (size_t)bt * R * C;
  if (VEC) {
    const int C4 = C >> 2;
    const float4 *src = reinterpret_cast<const float4 *>(am_row);
    for (int c = lane; c < C4; c += 32) {
      const float4 v = __ldg(src + c);
      for (int i = 0; i < R; ++i) st_stream_f4(reinterpret_cast<float4 *>(am_out + (size_t)i * C) + c, v);
    }
    for (int i = 0; i < R; ++i) {
      const int s = rg[i];
      const bool ok = (s >= 0 && s < S1);
      const float4 *lsrc = reinterpret_cast<const float4 *>(lm + ((size_t)b * S1 + (ok ? s : 0)) * C);
      float4 *ldst = reinterpret_cast<float4 *>(lm_out + (size_t)i * C);
      for (int c = lane; c < C4; c += 32) {
        float4 v = ok ? __ldg(lsrc + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        st_stream_f4(ldst + c, v);
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float v = am_row[c];
      for (int i = 0; i < R; ++i) am_out[(size_t)i * C + c] = v;
    }
    for (int i = 0; i < R; ++i) {
      const int s = rg[i];
      const bool ok = (s >= 0 && s < S1);
      const float *lsrc = lm + ((size_t)b * S1 + (ok ? s : 0)) * C;
      for (int c = lane; c < C; c += 32) lm_out[(size_t)i * C + c] = ok ? lsrc[c] : 0.f;
    }
  }
}

// A6 backward, am side: am_grad[b,t,:] = sum_i am_pruned_grad[b,t,i,:]
__global__ void __launch_bounds__(256) do_pruning_bwd_am_kernel(const float *am_p_grad, int BT, int R, int C,
                                                                float *am_grad) {
  const int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bt >= BT) return;
  const float *src = am_p_grad + (size_t)bt * R * C;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int i = 0; i < R; ++i) acc += src[(size_t)i * C + c];
    am_grad[(size_t)bt * C + c] = acc;
  }
}
// vector path: one CTA of 128 threads per (b,t), thread <-> one float4 column, the R rows loaded before the sum
template <int RMAX>
__global__ void __launch_bounds__(128) do_pruning_bwd_am_vec_kernel(const float *am_p_grad, int R, int C4,
                                                                    float *am_grad) {
  const int bt = blockIdx.x;
  const float4 *src = reinterpret_cast<const float4 *>(am_p_grad) + (size_t)bt * R * C4;
  float4 *dst = reinterpret_cast<float4 *>(am_grad) + (size_t)bt * C4;
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float4 v[RMAX];
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) v[i] = ld_stream_f4(src + (size_t)i * C4 + c);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
    dst[c] = acc;
  }
}

// A6 backward, lm side: lm_grad[b,s,:] = sum over (t,i) with ranges[b,t,i]==s  (the scatter-add TF autodiff
// derives for the gather of rnnt_loss.py:807-811, as a gather: deterministic, no atomics).  One CTA per (b,s):
//   1. the T x R indices of the utterance are tested once, cooperatively (any index pattern: hand-made,
//      repeated or non-consecutive ranges included): every thread takes a contiguous run of frames, counts its
//      hits, a block-wide exclusive scan gives its slot, and it writes its hits - the list is in (t,i) order;
//   2. every thread sums its float4 column over the list, four loads in flight.
// Shapes the list does not fit (or C % 4 != 0) walk the indices per element (LIST = false).
constexpr int kBwdLmThreads = 512;       // four groups of 128 column threads; the groups split the hit list
template <bool LIST>
__global__ void __launch_bounds__(kBwdLmThreads) do_pruning_bwd_lm_kernel(const float *lm_p_grad, const int32_t *ranges,
                                                                          int B, int S1, int T, int R, int C,
                                                                          float *lm_grad) {
  extern __shared__ int32_t hit_list[];       // [T * R] row offsets t * R + i, then (16-byte aligned) the partial sums
  __shared__ int warp_tot[kBwdLmThreads / 32];
  const int bs = blockIdx.x;
  const int b = bs / S1, s = bs - b * S1;
  const int32_t *rg = ranges + (size_t)b * T * R;
  const float *src = lm_p_grad + (size_t)b * T * R * C;
  if (LIST) {
    constexpr int NT = kBwdLmThreads, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int fpt = (T + NT - 1) / NT;                   // frames per thread
    const int t_lo = min(tid * fpt, T), t_hi = min(t_lo + fpt, T);
    int cnt = 0;
    for (int e = t_lo * R; e < t_hi * R; ++e) cnt += (rg[e] == s);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) { base += (j < w) ? warp_tot[j] : 0; total += warp_tot[j]; }
    int slot = base + incl - cnt;
    for (int e = t_lo * R; e < t_hi * R; ++e)
      if (rg[e] == s) hit_list[slot++] = e;
    __syncthreads();
    // A lattice row the band rests on for hundreds of frames has hundreds of hits (with an untrained model
    // nearly all frames of an utterance hit the same s_range rows): the thread groups take alternate runs of
    // 8 hits, 8 loads in flight per thread, so one CTA keeps 64 KB in flight.
    const int C4 = C >> 2;
    const int grp = tid >> 7, col = tid & 127;
    const float4 *src4 = reinterpret_cast<const float4 *>(src);
    float4 *dst = reinterpret_cast<float4 *>(lm_grad) + (size_t)bs * C4;
    constexpr int NG = NT / 128;
    float4 *partial = reinterpret_cast<float4 *>(hit_list + ((T * R + 3) & ~3));     // [NG - 1][128] sums of groups 1..
    for (int c0 = 0; c0 < C4; c0 += 128) {
      const int c = c0 + col;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C4) {
        for (int k0 = grp * 8; k0 < total; k0 += 8 * NG) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[j] = (k0 + j < total) ? ld_stream_f4(src4 + (size_t)hit_list[k0 + j] * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
      }
      if (grp > 0) partial[(grp - 1) * 128 + col] = acc;
      __syncthreads();
      if (grp == 0 && c < C4) {
#pragma unroll
        for (int g = 0; g < NG - 1; ++g) {
          const float4 o = partial[g * 128 + col];
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        dst[c] = acc;
      }
      __syncthreads();
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc = 0.f;
      for (int e = 0; e < T * R; ++e)
        if (rg[e] == s) acc += src[(size_t)e * C + c];
      lm_grad[(size_t)bs * C + c] = acc;
    }
  }
}

// (f2) fused additive joiner: logits[b,t,i,:] = am[b,t,:] + lm[b,ranges[b,t,i],:]
template <typename OutT>
__global__ void __launch_bounds__(256) pruned_add_joiner_kernel(const float *am, const float *lm,
                                                                const int32_t *ranges, int BT, int T, int S1,
                                                                int R, int C, OutT *logits) {
  const int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bt >= BT) return;
  const int b = bt / T;
  const float *am_row = am + (size_t)bt * C;
  for (int i = 0; i < R; ++i) {
    const int s = ranges[(size_t)bt * R + i];
    const bool ok = (s >= 0 && s < S1);
    const float *lsrc = lm + ((size_t)b * S1 + (ok ? s : 0)) * C;
    OutT *dst = logits + ((size_t)bt * R + i) * C;
    for (int c = lane; c < C; c += 32) {
      const float v = am_row[c] + (ok ? lsrc[c] : 0.f);
      dst[c] = static_cast<OutT>(v);
    }
  }
}

// float32 vector path of the fused joiner: one CTA of 128 threads per (b,t), thread <-> one
// float4 column, the am row and all R lm rows loaded before the R streaming stores.
template <int RMAX>
__global__ void __launch_bounds__(128) pruned_add_joiner_vec_kernel(const float *am, const float *lm,
                                                                    const int32_t *ranges, int T, int S1, int R,
                                                                    int C4, float *logits) {
  const int bt = blockIdx.x;
  const int b = bt / T;
  const int32_t *rg = ranges + (size_t)bt * R;
  const float4 *am_row = reinterpret_cast<const float4 *>(am) + (size_t)bt * C4;
  float4 *out = reinterpret_cast<float4 *>(logits) + (size_t)bt * R * C4;
  const float4 *lm_b = reinterpret_cast<const float4 *>(lm) + (size_t)b * S1 * C4;
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    const float4 a = __ldg(am_row + c);
    float4 l[RMAX];
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) {
        const int s = rg[i];
        l[i] = (s >= 0 && s < S1) ? __ldg(lm_b + (size_t)s * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) st_stream_f4(out + (size_t)i * C4 + c, make_float4(a.x + l[i].x, a.y + l[i].y, a.z + l[i].z, a.w + l[i].w));
    }
  }
}

// A6, am half on the copy engine: am_pruned[b,t,i,:] = am[b,t,:] does not depend on the ranges
// (rnnt_loss.py:802-806 broadcasts am before lm is gathered), so a caller may run it on a second stream
// beside the dependency-chain-bound kernels of the simple loss (normaliser: 128 CTAs, lattice recursion: 64
// CTAs on 148 SMs).  For that it must not take issue slots or SMs from them: a persistent grid of a few
// single-warp CTAs whose only instructions are 1-D bulk async copies (TMA engine; global -> shared ring ->
// R x global), and a shared-memory footprint (kBcStages x kBcStageBytes) that keeps the 204-224 KB CTAs of
// those kernels off the SMs it sits on instead of squeezing in beside them.
constexpr int kBcStages = 4, kBcLook = 2;
constexpr uint32_t kBcStageBytes = 32 * 1024;
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__global__ void __launch_bounds__(32) broadcast_am_kernel(const float *am, float *am_p, int BT, int R, int C,
                                                          int rows_per_chunk) {
  extern __shared__ __align__(128) unsigned char bc_smem[];
  __shared__ uint64_t bars[kBcStages];
  const int lane = threadIdx.x;
  const uint32_t row_bytes = (uint32_t)C * sizeof(float);
  const int nchunk_all = (BT + rows_per_chunk - 1) / rows_per_chunk;
  const int n = (nchunk_all - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // chunks of this CTA
  if (lane == 0) {
    for (int s = 0; s < kBcStages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  // lane 0 requests the chunks; ALL lanes issue the R x rows stores of a chunk (one thread issuing them was the
  // limit: 58 GB/s per CTA) - bulk async-groups are per thread, so every lane commits and waits for its own
  auto load = [&](int j) {
    const int chunk = blockIdx.x + j * gridDim.x, row0 = chunk * rows_per_chunk;
    const uint32_t bytes = (uint32_t)min(rows_per_chunk, BT - row0) * row_bytes;
    const int s = j % kBcStages;
    mbar_arrive_expect_tx(&bars[s], bytes);
    bulk_g2s(bc_smem + (size_t)s * kBcStageBytes, am + (size_t)row0 * C, bytes, &bars[s]);
  };
  if (lane == 0)
    for (int j = 0; j < kBcLook && j < n; ++j) load(j);
  for (int k = 0; k < n; ++k) {
    const int j = k + kBcLook;
    if (j < n) {                                  // warp-uniform
      // stage j % kBcStages was last read by the stores of chunk j - kBcStages: at most
      // kBcStages - kBcLook - 1 younger store groups (of any lane) may still be reading
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBcStages - kBcLook - 1) : "memory");
      __syncwarp();
      if (lane == 0) load(j);
    }
    const int s = k % kBcStages;
    mbar_wait(&bars[s], (uint32_t)((k / kBcStages) & 1));
    const int chunk = blockIdx.x + k * gridDim.x, row0 = chunk * rows_per_chunk;
    const int rows = min(rows_per_chunk, BT - row0);
    const unsigned char *src = bc_smem + (size_t)s * kBcStageBytes;
    for (int q = lane; q < rows * R; q += 32) {
      const int r = q / R, i = q - r * R;
      bulk_s2g(am_p + ((size_t)(row0 + r) * R + i) * C, src + (size_t)r * row_bytes, row_bytes);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// bfloat16 output of the fused joiner (BASELINE configs[3]: bf16 joiner logits): thread <-> 8 columns, two
// float4 of the am row and of every lm row in, one 16-byte store of 8 bf16 out
template <int RMAX>
__global__ void __launch_bounds__(128) pruned_add_joiner_vec_bf16_kernel(const float *am, const float *lm,
                                                                         const int32_t *ranges, int T, int S1, int R,
                                                                         int C8, __nv_bfloat16 *logits) {
  const int bt = blockIdx.x;
  const int b = bt / T;
  const int32_t *rg = ranges + (size_t)bt * R;
  const float4 *am_row = reinterpret_cast<const float4 *>(am) + (size_t)bt * C8 * 2;
  uint4 *out = reinterpret_cast<uint4 *>(logits) + (size_t)bt * R * C8;
  const float4 *lm_b = reinterpret_cast<const float4 *>(lm) + (size_t)b * S1 * C8 * 2;
  for (int c = threadIdx.x; c < C8; c += blockDim.x) {
    const float4 a0 = __ldg(am_row + 2 * c), a1 = __ldg(am_row + 2 * c + 1);
    float4 l0[RMAX], l1[RMAX];
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) {
        const int s = rg[i];
        const bool ok = s >= 0 && s < S1;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        l0[i] = ok ? __ldg(lm_b + ((size_t)s * C8 + c) * 2) : z;
        l1[i] = ok ? __ldg(lm_b + ((size_t)s * C8 + c) * 2 + 1) : z;
      }
    }
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(a0.x + l0[i].x, a0.y + l0[i].y);
        const __nv_bfloat162 p1 = __floats2bfloat162_rn(a0.z + l0[i].z, a0.w + l0[i].w);
        const __nv_bfloat162 p2 = __floats2bfloat162_rn(a1.x + l1[i].x, a1.y + l1[i].y);
        const __nv_bfloat162 p3 = __floats2bfloat162_rn(a1.z + l1[i].z, a1.w + l1[i].w);
        uint4 v;
        v.x = *reinterpret_cast<const uint32_t *>(&p0); v.y = *reinterpret_cast<const uint32_t *>(&p1);
        v.z = *reinterpret_cast<const uint32_t *>(&p2); v.w = *reinterpret_cast<const uint32_t *>(&p3);
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + (size_t)i * C8 + c), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
int launch_cummin(const int32_t *in, int32_t *out, int rows, int n, cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return FRN_OK;
  count_launch(), cummin_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(in, out, rows, n);
  return check_launch();
}

int launch_prune_ranges(const float *px_grad, const float *py_grad, const int32_t *boundary, int B, int S, int T,
                        int T1, int R, int32_t *ranges, int32_t *s_begin_ws, cudaStream_t stream) {
  const int r_fix = (T1 == T) ? 2 : R;  // rnnt_loss.py:756
  const size_t smem = (size_t)(T + 16) * sizeof(int32_t);
  if (smem > 200 * 1024) return FRN_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(prune_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return note_cuda_error(e);
  }
  // columns per block: as many as keep the tile under ~100 KB (two blocks per SM)
  int cols = kArgmaxThreads;
  while (cols > 8 && (size_t)(2 * S + 1) * cols * sizeof(float) > 100 * 1024) cols >>= 1;
  // a full warp of columns (128-byte rows) if 2 x kArgmaxTileBytes still fit an SM: S = 400 (c4, c5) gets 32, not 16
  if (cols < 32 && (size_t)(2 * S + 1) * 32 * sizeof(float) <= kArgmaxTileBytes) cols = 32;
  const size_t tile_bytes = (size_t)(2 * S + 1) * cols * sizeof(float);
  if (tile_bytes > 200 * 1024) return FRN_EUNSUPPORTED;
  dim3 grid((T + cols - 1) / cols, B);
  if (tile_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(prune_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
    if (e != cudaSuccess) return note_cuda_error(e);
  }
  count_launch(), prune_argmax_kernel<<<grid, cols, tile_bytes, stream>>>(px_grad, py_grad, boundary, S, T, T1, R, s_begin_ws);
  int rc = check_launch();
  if (rc) return rc;
  count_launch(), prune_fixup_kernel<<<B, kPruneThreads, smem, stream>>>(s_begin_ws, T, R, r_fix, ranges);
  return check_launch();
}

int launch_do_pruning(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                      float *am_p, float *lm_p, cudaStream_t stream) {
  const int BT = B * T;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) |
                                     reinterpret_cast<uintptr_t>(am_p) | reinterpret_cast<uintptr_t>(lm_p)) % 16 == 0);
  // one half only (the am broadcast does not depend on the ranges, callers may run it early on another stream)
  if (!am_p || !lm_p) {
    if (!(vec && R <= 8)) return FRN_EUNSUPPORTED;
    if (am_p) count_launch(), do_pruning_vec_kernel<8, true, false><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4, am_p, lm_p);
    else count_launch(), do_pruning_vec_kernel<8, false, true><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4, am_p, lm_p);
    return check_launch();
  }
  if (vec && R <= 8) count_launch(), do_pruning_vec_kernel<8, true, true><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4, am_p, lm_p);
  else if (vec) count_launch(), do_pruning_kernel<true><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C, am_p, lm_p);
  else count_launch(), do_pruning_kernel<false><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C, am_p, lm_p);
  return check_launch();
}

int launch_broadcast_am(const float *am, int B, int T, int R, int C, float *am_p, int max_ctas, cudaStream_t stream) {
  const int BT = B * T;
  const size_t row_bytes = (size_t)C * sizeof(float);
  if (C % 4 != 0 || row_bytes > kBcStageBytes ||
      ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(am_p)) % 16) != 0)
    return FRN_EUNSUPPORTED;     // bulk copies move multiples of 16 bytes between 16-byte aligned addresses
  const int rows = (int)(kBcStageBytes / row_bytes);
  const int nchunk = (BT + rows - 1) / rows;
  const int grid = std::max(1, std::min(max_ctas > 0 ? max_ctas : 20, nchunk));
  const size_t smem = (size_t)kBcStages * kBcStageBytes;
  cudaError_t e = cudaFuncSetAttribute(broadcast_am_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return note_cuda_error(e);
  count_launch(), broadcast_am_kernel<<<grid, 32, smem, stream>>>(am, am_p, BT, R, C, rows);
  return check_launch();
}

int launch_do_pruning_add(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                          float *am_p, float *lm_p, float *logits, cudaStream_t stream) {
  const int BT = B * T;
  if (!am_p) {     // am_pruned is written elsewhere (frn_broadcast_am_pruned): lm_pruned and the sum only
    const bool v = (C % 4 == 0) && R <= 8 &&
                   ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) | reinterpret_cast<uintptr_t>(lm_p) |
                     reinterpret_cast<uintptr_t>(logits)) % 16 == 0);
    if (!v) return FRN_EUNSUPPORTED;
    count_launch(), do_pruning_vec_kernel<8, false, true, true><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4, am_p, lm_p, logits);
    return check_launch();
  }
  const bool vec = (C % 4 == 0) && R <= 8 &&
                   ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) | reinterpret_cast<uintptr_t>(am_p) |
                     reinterpret_cast<uintptr_t>(lm_p) | reinterpret_cast<uintptr_t>(logits)) % 16 == 0);
  if (!vec) {   // general shapes: the two separate passes
    int rc = launch_do_pruning(am, lm, ranges, B, S, T, R, C, am_p, lm_p, stream);
    if (rc) return rc;
    return launch_add(am_p, lm_p, logits, (size_t)BT * R * C, stream);
  }
  count_launch(), do_pruning_vec_kernel<8, true, true, true><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4, am_p, lm_p, logits);
  return check_launch();
}

int launch_do_pruning_bwd(const float *am_p_grad, const float *lm_p_grad, const int32_t *ranges, int B, int S,
                          int T, int R, int C, float *am_grad, float *lm_grad, cudaStream_t stream) {
  const int BT = B * T;
  auto aligned16 = [](const void *a, const void *b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
  };
  if (am_grad) {
    if (C % 4 == 0 && R <= 8 && aligned16(am_p_grad, am_grad))
      count_launch(), do_pruning_bwd_am_vec_kernel<8><<<BT, 128, 0, stream>>>(am_p_grad, R, C / 4, am_grad);
    else
      count_launch(), do_pruning_bwd_am_kernel<<<(BT + 7) / 8, 256, 0, stream>>>(am_p_grad, BT, R, C, am_grad);
    int rc = check_launch();
    if (rc) return rc;
  }
  if (lm_grad) {
    const size_t list_bytes = (((size_t)T * R + 3) & ~(size_t)3) * sizeof(int32_t) +
                              (kBwdLmThreads / 128 - 1) * 128 * sizeof(float4);
    if (C % 4 == 0 && aligned16(lm_p_grad, lm_grad) && list_bytes <= 160 * 1024) {
      if (list_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(do_pruning_bwd_lm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)list_bytes);
        if (e != cudaSuccess) return note_cuda_error(e);
      }
      count_launch(), do_pruning_bwd_lm_kernel<true><<<B * (S + 1), kBwdLmThreads, list_bytes, stream>>>(
          lm_p_grad, ranges, B, S + 1, T, R, C, lm_grad);
    } else {
      count_launch(), do_pruning_bwd_lm_kernel<false><<<B * (S + 1), kBwdLmThreads, 0, stream>>>(lm_p_grad, ranges, B, S + 1,
                                                                                              T, R, C, lm_grad);
    }
    return check_launch();
  }
  return FRN_OK;
}

int launch_pruned_add_joiner(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R,
                             int C, int out_dtype, void *logits, cudaStream_t stream) {
  const int BT = B * T;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) |
                                     reinterpret_cast<uintptr_t>(logits)) % 16 == 0);
  if (out_dtype == FRN_F32 && vec && R <= 8) {
    count_launch(), pruned_add_joiner_vec_kernel<8><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 4,
                                                                         static_cast<float *>(logits));
    return check_launch();
  }
  if (out_dtype == FRN_BF16 && C % 8 == 0 && R <= 8 &&
      ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) | reinterpret_cast<uintptr_t>(logits)) % 16 == 0)) {
    count_launch(), pruned_add_joiner_vec_bf16_kernel<8><<<BT, 128, 0, stream>>>(am, lm, ranges, T, S + 1, R, C / 8,
                                                                              static_cast<__nv_bfloat16 *>(logits));
    return check_launch();
  }
  if (out_dtype == FRN_F32)
    count_launch(), pruned_add_joiner_kernel<float><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C,
                                                                      static_cast<float *>(logits));
  else if (out_dtype == FRN_BF16)
    count_launch(), pruned_add_joiner_kernel<__nv_bfloat16><<<(BT + 7) / 8, 256, 0, stream>>>(
        am, lm, ranges, BT, T, S + 1, R, C, static_cast<__nv_bfloat16 *>(logits));
  else return FRN_EINVAL;
  return check_launch();
}

}  // namespace frn
