// Prune ranges (A5), cummin (op "Cummin"), pruning gather (A6) and its gradient.
//
// Replaces get_rnnt_prune_ranges / _adjust_pruning_lower_bound /
// _monotonic_lower_bound (rnnt_loss.py:553-761: ~15 TF kernels + two calls of
// the reference's scan kernel, mutual_information_cuda.cu:895-1012) and
// do_rnnt_pruning (rnnt_loss.py:763-812).  Integer results are bit-exact by
// construction: the float part walks s sequentially per (b,t) column, i.e. the
// summation order of a sequential cumsum, with first-index
// argmax; the int32 fix-ups along t are order independent.
#include "common.cuh"

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
// A5 step 1: s_begin[b,t] = argmax_k ( sum_{k<=s<k+R} py_grad[s,t] - px_grad[k-1,t] )
// (rnnt_loss.py:722-748), thread per (b,t) column, coalesced along t.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prune_argmax_kernel(const float *px_grad, const float *py_grad,
                                                           const int32_t *boundary, int S, int T, int T1,
                                                           int R, int32_t *s_begin) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int s_end = boundary[4 * b + 2], t_end = boundary[4 * b + 3];
  const int S1 = S + 1;
  int best_k = 0;
  if (t < t_end - 1) {
    const float *py = py_grad + (size_t)b * S1 * T + t;
    const float *px = px_grad + (size_t)b * S * T1 + t;
    const int nk = S1 - R + 1;
    float cs_lo = 0.f, cs_hi = 0.f;
    for (int s = 0; s < R; ++s) cs_hi = cs_hi + py[(size_t)s * T];  // cs[R], sequential
    float best = 0.f;
    for (int k = 0; k < nk; ++k) {
      float fin = cs_hi - cs_lo;
      if (k > 0) fin = fin - px[(size_t)(k - 1) * T1];
      if (k == 0 || fin > best) { best = fin; best_k = k; }
      cs_lo = cs_lo + py[(size_t)k * T];
      if (k + R < S1) cs_hi = cs_hi + py[(size_t)(k + R) * T];
    }
  } else {
    best_k = max(s_end - R + 1, 0);  // padding frames, rnnt_loss.py:744-748
  }
  s_begin[(size_t)b * T + t] = best_k;
}

// ---------------------------------------------------------------------------
// A5 step 2: monotonic fix-ups (rnnt_loss.py:623-641) + range expansion
// (:758-759).  One warp per utterance; reverse running minima by chunked warp
// scans walking t from the end.
// ---------------------------------------------------------------------------
__device__ void rev_cummin_row(int32_t *x, int T, int lane) {
  int carry = INT32_MAX;
  for (int base = T - 1; base >= 0; base -= 32) {
    const int i = base - lane;  // lane 0 holds the largest t
    int v = (i >= 0) ? x[i] : INT32_MAX;
    v = min(warp_incl_min_scan(v, lane), carry);
    if (i >= 0) x[i] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}

__global__ void __launch_bounds__(32) prune_fixup_kernel(int32_t *s_begin, int T, int r, int R, int32_t *ranges) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int32_t *x = s_begin + (size_t)b * T;
  rev_cummin_row(x, T, lane);
  __syncwarp();
  for (int t = lane; t < T; t += 32) x[t] = -(x[t] - (r - 1) * t);
  __syncwarp();
  rev_cummin_row(x, T, lane);
  __syncwarp();
  for (int t = lane; t < T; t += 32) x[t] = -(max(x[t], 0) - (r - 1) * t);
  __syncwarp();
  int32_t *out = ranges + (size_t)b * T * R;
  for (int i = lane; i < T * R; i += 32) {
    const int t = i / R;
    out[i] = x[t] + (i - t * R);
  }
}

// ---------------------------------------------------------------------------
// A6 forward: am_pruned[b,t,i,:] = am[b,t,:], lm_pruned[b,t,i,:] = lm[b,ranges[b,t,i],:]
// One warp per (b,t); 128-bit accesses when C % 4 == 0 (rows are then 16-byte
// aligned).  Out-of-range indices produce zeros (tf.gather on GPU).
// ---------------------------------------------------------------------------
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

// A6 backward, lm side: lm_grad[b,s,:] = sum over (t,i) with ranges[b,t,i]==s.
// Gather formulation (deterministic, no atomics): one warp per (b,s) scans the
// frames; since ranges[b,t,i] = ranges[b,t,0] + i the hit test is a subtraction.
__global__ void __launch_bounds__(256) do_pruning_bwd_lm_kernel(const float *lm_p_grad, const int32_t *ranges,
                                                                int B, int S1, int T, int R, int C,
                                                                float *lm_grad) {
  const int bs = blockIdx.x;
  const int b = bs / S1, s = bs - b * S1;
  const int32_t *rg = ranges + (size_t)b * T * R;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      for (int i = 0; i < R; ++i)
        if (rg[(size_t)t * R + i] == s) acc += lm_p_grad[(((size_t)b * T + t) * R + i) * C + c];
    }
    lm_grad[(size_t)bs * C + c] = acc;
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

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
int launch_cummin(const int32_t *in, int32_t *out, int rows, int n, cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return FRN_OK;
  cummin_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(in, out, rows, n);
  return check_launch();
}

int launch_prune_ranges(const float *px_grad, const float *py_grad, const int32_t *boundary, int B, int S, int T,
                        int T1, int R, int32_t *ranges, int32_t *s_begin_ws, cudaStream_t stream) {
  dim3 grid((T + 127) / 128, B);
  prune_argmax_kernel<<<grid, 128, 0, stream>>>(px_grad, py_grad, boundary, S, T, T1, R, s_begin_ws);
  int rc = check_launch();
  if (rc) return rc;
  const int r = (T1 == T) ? 2 : R;  // rnnt_loss.py:756
  prune_fixup_kernel<<<B, 32, 0, stream>>>(s_begin_ws, T, r, R, ranges);
  return check_launch();
}

int launch_do_pruning(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R, int C,
                      float *am_p, float *lm_p, cudaStream_t stream) {
  const int BT = B * T;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm) |
                                     reinterpret_cast<uintptr_t>(am_p) | reinterpret_cast<uintptr_t>(lm_p)) % 16 == 0);
  if (vec) do_pruning_kernel<true><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C, am_p, lm_p);
  else do_pruning_kernel<false><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C, am_p, lm_p);
  return check_launch();
}

int launch_do_pruning_bwd(const float *am_p_grad, const float *lm_p_grad, const int32_t *ranges, int B, int S,
                          int T, int R, int C, float *am_grad, float *lm_grad, cudaStream_t stream) {
  const int BT = B * T;
  if (am_grad) {
    do_pruning_bwd_am_kernel<<<(BT + 7) / 8, 256, 0, stream>>>(am_p_grad, BT, R, C, am_grad);
    int rc = check_launch();
    if (rc) return rc;
  }
  if (lm_grad) {
    do_pruning_bwd_lm_kernel<<<B * (S + 1), 256, 0, stream>>>(lm_p_grad, ranges, B, S + 1, T, R, C, lm_grad);
    return check_launch();
  }
  return FRN_OK;
}

int launch_pruned_add_joiner(const float *am, const float *lm, const int32_t *ranges, int B, int S, int T, int R,
                             int C, int out_dtype, void *logits, cudaStream_t stream) {
  const int BT = B * T;
  if (out_dtype == FRN_F32)
    pruned_add_joiner_kernel<float><<<(BT + 7) / 8, 256, 0, stream>>>(am, lm, ranges, BT, T, S + 1, R, C,
                                                                      static_cast<float *>(logits));
  else if (out_dtype == FRN_BF16)
    pruned_add_joiner_kernel<__nv_bfloat16><<<(BT + 7) / 8, 256, 0, stream>>>(
        am, lm, ranges, BT, T, S + 1, R, C, static_cast<__nv_bfloat16 *>(logits));
  else return FRN_EINVAL;
  return check_launch();
}

}  // namespace frn
