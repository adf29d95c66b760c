// Pruned log-probs (A7), band <-> lattice glue for the fused pruned loss (A8)
// and the logits gradient.
//
// Replaces get_rnnt_logprobs_pruned (rnnt_loss.py:853-1020: reduce_logsumexp +
// ~12 gather/concat/roll/transpose kernels on lattice-sized tensors) and the
// TensorFlow autodiff chain back to the joiner logits.  The fused loss never
// builds the dense [B,S,T+1] lattice in the reference layout: band entries go
// straight into the diagonal-major planes the chain kernel consumes
// (equivalence: SURVEY.md §8a-A8), and occupation counts come back compact
// ([B,T,R]) to feed the logits-gradient kernel.
#include "common.cuh"

namespace frn {

template <typename T> struct Vec8;  // 16-byte vector loads of the logits
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

template <typename T> constexpr int kVecElems = 16 / sizeof(T);

// band position of entry i at frame t: the reference rolls by ranges[b,t,0]
// modulo S+1 (rnnt_loss.py:849)
__device__ __forceinline__ int band_row(int r0, int i, int S1) {
  int s = r0 + i;
  s %= S1;
  if (s < 0) s += S1;
  return s;
}

// ---------------------------------------------------------------------------
// A7 forward: one warp per (b,t,i) row of C logits, single pass online
// log-sum-exp, symbol/blank gather.  pxc = l[sym'] - lse, pyc = l[blank] - lse.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pruned_lse_kernel(const T *logits, const int32_t *symbols,
                                                         const int32_t *ranges, int BTR, int TR, int S, int R,
                                                         int C, int term, float *pxc, float *pyc, float *lse_out) {
  constexpr int V = kVecElems<T>;
  constexpr int kU = 4;   // 16-byte loads issued back to back per lane before any arithmetic
  constexpr uint32_t kPad = sizeof(T) == 4 ? 0xff800000u : 0xff80ff80u;   // -inf as float32 / as two bfloat16
  const int lane = threadIdx.x & 31;
  const int nv = C / V;
  // fast path: rows are 16-byte aligned and fit one batch of kU vector loads per lane
  const bool vec = (C % V == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15u) == 0);
  const bool one_batch = vec && nv <= 32 * kU;

  auto load_batch = [&](const uint4 *p, int cb, uint4 (&raw)[kU]) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int c = cb + u * 32 + lane;
      raw[u] = make_uint4(kPad, kPad, kPad, kPad);      // -inf in every element: absent vectors drop out of max / sum
      if (c < nv)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw[u].x), "=r"(raw[u].y), "=r"(raw[u].z), "=r"(raw[u].w) : "l"(p + c));
    }
  };
  // symbol context of a band entry (rnnt_loss.py:954-958): -1 = no symbol arc
  auto load_sym = [&](int row, bool &s_ok) {
    int s, sym = -1;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(s) : "l"(ranges + row));
    s_ok = (s >= 0 && s <= S);
    if (s_ok) {
      if (s < S) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(sym) : "l"(symbols + (size_t)(gridDim.y - 1 - blockIdx.y) * S + s));
      else sym = term;
    }
    return (sym < 0 || sym >= C) ? -1 : sym;
  };

  // One row per warp (a persistent grid-stride variant with next-row prefetch measured 1.5x
  // slower: fewer, longer-lived warps hide the memory latency worse than many short ones).  The
  // row's vector loads are requested before the dependent index loads so that all are in flight
  // together; the two gathered logits are two more scalar loads issued as soon as their index is
  // known (they hit L2 behind the row itself), which keeps per-element work to max / ex2 / add.
  // grid: x = groups of 8 rows inside an utterance, y = utterance (no integer division by T*R)
  // Rows are taken LAST FIRST: the kernel that produced the logits (the joiner) wrote them first to last, so
  // what is still in the 126 MB L2 is their tail; and this kernel then leaves the head there for the
  // gradient kernel, which runs first to last (measured: 105.5 -> 103.6 us for the pruned loss at c2).
  const int by = gridDim.y - 1 - blockIdx.y;
  const int local = (gridDim.x - 1 - blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (local >= TR) return;
  const int row = by * TR + local;
  const T *src = logits + (size_t)row * C;
  uint4 raw[kU];
  if (one_batch) load_batch(reinterpret_cast<const uint4 *>(src), 0, raw);
  float g_term = to_f(src[term]);
  bool s_ok;
  const int sym = load_sym(row, s_ok);
  float g_sym = (sym >= 0) ? to_f(src[sym]) : 0.f;
  float m = -INFINITY, ssum = 0.f;     // running max and sum of 2^((x - m) log2e)
  if (one_batch) {
    float x[kU][V];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const T *e = reinterpret_cast<const T *>(&raw[u]);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        x[u][j] = to_f(e[j]);
        m = fmaxf(m, x[u][j]);
      }
    }
    m = warp_max(m);                   // one max for the whole row: no rescaling afterwards
    if (m > -INFINITY) {
      const float m2 = m * kLog2e;
#pragma unroll
      for (int u = 0; u < kU; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) ssum += ex2_approx(fmaf(x[u][j], kLog2e, -m2));
    }
  } else if (vec) {
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    for (int cb = 0; cb < nv; cb += 32 * kU) {
      load_batch(p, cb, raw);
      float x[kU][V];
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const T *e = reinterpret_cast<const T *>(&raw[u]);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          x[u][j] = to_f(e[j]);
          mx = fmaxf(mx, x[u][j]);
        }
      }
      const float mn = fmaxf(m, mx);
      if (mn > -INFINITY) {
        const float mn2 = mn * kLog2e;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
          for (int j = 0; j < V; ++j) acc += ex2_approx(fmaf(x[u][j], kLog2e, -mn2));
        ssum = ssum * ex2_approx((m - mn) * kLog2e) + acc;
        m = mn;
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float x = to_f(src[c]);
      const float mn = fmaxf(m, x);
      if (mn > -INFINITY) {
        ssum = ssum * ex2_approx((m - mn) * kLog2e) + ex2_approx((x - mn) * kLog2e);
        m = mn;
      }
    }
  }
  // combine lanes
  const float M = warp_max(m);
  const float scaled = (m == -INFINITY) ? 0.f : ssum * ex2_approx((m - M) * kLog2e);
  const float tot = warp_sum(scaled);
  const float lse = (M == -INFINITY) ? -INFINITY : M + logf(tot);
  if (lane == 0) {
    float vx = -INFINITY, vy = -INFINITY;
    if (s_ok) {
      if (sym >= 0) vx = g_sym - lse;
      vy = g_term - lse;
    }
    pxc[row] = vx; pyc[row] = vy; lse_out[row] = lse;
  }
}

// ---------------------------------------------------------------------------
// band -> diagonal-major planes (after the planes were filled with the
// sentinel).  Thread per (b,t,i).
// ---------------------------------------------------------------------------
struct BandParams {
  const float *pxc, *pyc;   // [B][T][R]
  const int32_t *ranges;    // [B][T][R]
  const int32_t *boundary;
  int S, T, R, P, Dn, k, rnnt_type;
  float delay_penalty;
};

__global__ void __launch_bounds__(256) fill_dead_arcs_kernel(float4 *XY, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float dead = __int_as_float(kNegI);
  if (i < n) XY[i] = make_float4(0.f, dead, 0.f, dead);
}

__global__ void __launch_bounds__(256) skew_band_kernel(BandParams p, float4 *XY, int BTR) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BTR) return;
  const int TR = p.T * p.R;
  const int b = idx / TR, rem = idx - b * TR, t = rem / p.R, i = rem - t * p.R;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, s_end = bd.z, t_end = bd.w;
  if (s_end < s_begin || t_end < t_begin || s_begin < 0 || t_begin < 0 || s_end > p.S || t_end > p.T) return;
  if (t < t_begin || t >= t_end) return;  // arcs leave frames t_begin..t_end-1 only
  const int r0 = p.ranges[(size_t)(b * p.T + t) * p.R];
  const int s = band_row(r0, i, p.S + 1);
  if (s < s_begin || s > s_end) return;
  const int sp = s - s_begin, tp = t - t_begin;
  const size_t plane = (size_t)b * p.Dn * p.P;
  const int noff = p.k ? 0 : 1;
  // py arc (s,t) -> (s,t+1): destination diagonal (tp+1) + k*sp
  {
    const float v = p.pyc[idx];
    float2 *cell = reinterpret_cast<float2 *>(XY + plane + (size_t)(tp + 1 + p.k * sp) * p.P + sp);
    cell[1] = encode_arc(fmaxf(v * kLog2e, kNeg));
  }
  // px arc (s,t) -> (s+1,t+noff)
  if (s < p.S && s < s_end) {
    float v = p.pxc[idx];
    if (p.rnnt_type == FRN_CONSTRAINED) {
      // px += py[s+1][t] (rnnt_loss.py:1018); py[s+1][t] is band entry i+1
      float add = -INFINITY;
      if (i + 1 < p.R && band_row(r0, i + 1, p.S + 1) == s + 1) add = p.pyc[idx + 1];
      v += add;
    }
    if (p.delay_penalty != 0.f) v += delay_penalty_value(t_end, t, p.delay_penalty);
    float2 *cell = reinterpret_cast<float2 *>(XY + plane + (size_t)(tp + noff + p.k * (sp + 1)) * p.P + sp + 1);
    cell[0] = encode_arc(fmaxf(v * kLog2e, kNeg));
  }
}

// ---------------------------------------------------------------------------
// compact occupation counts of the band arcs + scores.  Thread per (b,t,i).
// gxc/gyc[b,t,i]: d score_b / d pxc[b,t,i], d score_b / d pyc[b,t,i]
// (constrained: the px arc's count is folded into the py entry it borrowed).
// ---------------------------------------------------------------------------
struct BandFinalizeParams {
  const float2 *A;
  const float4 *Bq;
  const int32_t *ranges, *boundary;
  int S, T, R, P, Dn, k, rnnt_type;
};

__device__ __forceinline__ bool bd_ok(const int4 &bd, int S, int T) {
  return bd.z >= bd.x && bd.w >= bd.y && bd.x >= 0 && bd.y >= 0 && bd.z <= S && bd.w <= T;
}

__global__ void __launch_bounds__(256) finalize_band_kernel(BandFinalizeParams p, float *gxc, float *gyc,
                                                            float *scores, int B) {
  const int TR = p.T * p.R;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B && scores) {
    const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * idx);
    float v = 0.f;
    if (bd_ok(bd, p.S, p.T)) {
      const int Sb = bd.z - bd.x, Tb = bd.w - bd.y, Db = Tb + p.k * Sb;
      v = lattice_score(p.A[((size_t)idx * p.Dn + Db) * p.P + Sb]);
    }
    scores[idx] = v;
  }
  if (idx >= B * TR || gxc == nullptr) return;
  const int b = idx / TR, rem = idx - b * TR, t = rem / p.R, i = rem - t * p.R;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  float vx = 0.f, vy = 0.f;
  if (bd_ok(bd, p.S, p.T) && t >= bd.y && t < bd.w) {
    const int Sb = bd.z - bd.x, Tb = bd.w - bd.y, Db = Tb + p.k * Sb;
    const size_t plane = (size_t)b * p.Dn * p.P;
    const float2 tot = p.A[plane + (size_t)Db * p.P + Sb];
    if (tot.x > 0.f) {
      const float inv_tot = 1.0f / tot.x;
      const int tot_o = __float_as_int(tot.y);
      const int r0 = p.ranges[(size_t)(b * p.T + t) * p.R];
      const int s = band_row(r0, i, p.S + 1);
      auto occ = [&](int sa, bool px_arc) -> float {
        // occupation of the px/py arc leaving (sa, t)
        if (sa < bd.x || sa > bd.z) return 0.f;
        const int sp = sa - bd.x, tp = t - bd.y;
        if (px_arc && !(sa < p.S && sp < Sb)) return 0.f;
        const int d = tp + p.k * sp;
        const size_t at = plane + (size_t)d * p.P + sp;
        const float4 bq = p.Bq[at];
        return (px_arc ? bq.x : bq.y) * occupation_scale(p.A[at], __float_as_int(bq.z), tot_o, inv_tot);
      };
      vx = occ(s, true);
      vy = occ(s, false);
      if (p.rnnt_type == FRN_CONSTRAINED && i >= 1 && band_row(r0, i - 1, p.S + 1) == s - 1)
        vy += occ(s - 1, true);
    }
  }
  gxc[idx] = vx;
  gyc[idx] = vy;
}

// ---------------------------------------------------------------------------
// A7 backward: dlogits[row,c] = g_b ( [c=sym'] gx + [c=blank] gy - softmax_c (gx+gy) )
// one warp per row, softmax recomputed from the stored log-sum-exp.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pruned_logits_grad_kernel(const T *logits, const int32_t *symbols,
                                                                 const int32_t *ranges, const float *lse,
                                                                 const float *gxc, const float *gyc,
                                                                 const float *scores_grad, int BTR, int TR,
                                                                 int S, int C, int term, T *dlogits) {
  constexpr int V = kVecElems<T>;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= BTR) return;
  const int b = row / TR;
  const float g = scores_grad ? scores_grad[b] : 1.f;
  const float gx = gxc[row] * g, gy = gyc[row] * g;
  const float gs = gx + gy;
  const float l2 = lse[row] * kLog2e;
  const int s = ranges[row];
  int sym = -1;
  if (s >= 0 && s <= S) sym = (s < S) ? symbols[(size_t)b * S + s] : term;
  const T *src = logits + (size_t)row * C;
  T *dst = dlogits + (size_t)row * C;
  const bool vec = (C % V == 0) && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0);
  if (vec) {
    const int nv = C / V;
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    uint4 *q = reinterpret_cast<uint4 *>(dst);
    // vectors holding the symbol / blank column get their fix-up in a (rarely taken) branch, so
    // the common vector costs multiply-add, ex2, multiply per element and nothing else
    const int vsym = (sym >= 0) ? sym / V : -1, vterm = term / V;
    const bool zero_row = (gs == 0.f);                 // warp-uniform: no softmax term, nothing to read
    for (int c = lane; c < nv; c += 32) {
      uint4 raw = make_uint4(0, 0, 0, 0), outv;
      if (!zero_row)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p + c));
      const T *e = reinterpret_cast<const T *>(&raw);
      T *o = reinterpret_cast<T *>(&outv);
      float d[V];
#pragma unroll
      for (int j = 0; j < V; ++j) d[j] = zero_row ? 0.f : -gs * ex2_approx(fmaf(to_f(e[j]), kLog2e, -l2));
      if (c == vsym || c == vterm) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int cc = c * V + j;
          if (cc == sym) d[j] += gx;
          if (cc == term) d[j] += gy;
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = from_f<T>(d[j]);
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(q + c), "r"(outv.x),
                   "r"(outv.y), "r"(outv.z), "r"(outv.w) : "memory");
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float d = 0.f;
      if (gs != 0.f) d = -gs * ex2_approx(to_f(src[c]) * kLog2e - l2);
      if (c == sym) d += gx;
      if (c == term) d += gy;
      dst[c] = from_f<T>(d);
    }
  }
}

// ---------------------------------------------------------------------------
// dense px/py in the reference layout from the compact band
// (public get_rnnt_logprobs_pruned).  Thread per output element.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) band_to_dense_kernel(const float *pxc, const float *pyc,
                                                            const int32_t *ranges, const int32_t *boundary,
                                                            int B, int S, int T, int T1, int R, int rnnt_type,
                                                            float *px, float *py) {
  const int S1 = S + 1;
  const size_t n_px = (size_t)B * S * T1, n_py = (size_t)B * S1 * T;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto band_val = [&](const float *src, int b, int s, int t) -> float {
    const int r0 = ranges[(size_t)(b * T + t) * R];
    int i = (s - r0) % S1;
    if (i < 0) i += S1;
    return (i < R) ? src[(size_t)(b * T + t) * R + i] : -INFINITY;
  };
  if (gid < n_px) {
    const int b = (int)(gid / ((size_t)S * T1));
    const int rem = (int)(gid - (size_t)b * S * T1);
    const int s = rem / T1, t = rem - s * T1;
    float v = -INFINITY;
    if (t < T) {
      v = band_val(pxc, b, s, t);
      if (rnnt_type == FRN_CONSTRAINED) v += band_val(pyc, b, s + 1, t);
      if (rnnt_type == FRN_REGULAR && t == boundary[4 * b + 3]) v = -INFINITY;  // fix_for_boundary
    }
    px[gid] = v;
  } else if (gid < n_px + n_py) {
    const size_t g2 = gid - n_px;
    const int b = (int)(g2 / ((size_t)S1 * T));
    const int rem = (int)(g2 - (size_t)b * S1 * T);
    const int s = rem / T, t = rem - s * T;
    py[g2] = band_val(pyc, b, s, t);
  }
}

// Transpose of band_to_dense_kernel: cotangents of the dense px / py -> cotangents of the band log-probs
// (backward of the public get_rnnt_logprobs_pruned; TF autodiff through rnnt_loss.py:968-1018).  Entries the
// forward overwrote with -inf (frame t_end of the regular lattice, :51-60) pass nothing back.
__global__ void __launch_bounds__(256) dense_to_band_kernel(const float *dpx, const float *dpy,
                                                            const int32_t *ranges, const int32_t *boundary,
                                                            int B, int S, int T, int T1, int R, int rnnt_type,
                                                            float *gxc, float *gyc) {
  const int S1 = S + 1;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B * T * R) return;
  const int bt = row / R, i = row - bt * R;
  const int b = bt / T, t = bt - b * T;
  int s = (ranges[(size_t)bt * R] + i) % S1;
  if (s < 0) s += S1;
  const bool x_dead = rnnt_type == FRN_REGULAR && t == boundary[4 * b + 3];
  float gx = 0.f, gy = dpy[((size_t)b * S1 + s) * T + t];
  if (s < S && !x_dead) gx = dpx[((size_t)b * S + s) * T1 + t];
  // constrained: px[s-1,t] = pxc[s-1,t] + pyc[s,t] (rnnt_loss.py:1015-1018); row s-1 is band entry i-1
  if (rnnt_type == FRN_CONSTRAINED && i >= 1 && s >= 1) gy += dpx[((size_t)b * S + s - 1) * T1 + t];
  gxc[row] = gx;
  gyc[row] = gy;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
int launch_pruned_lse(const void *logits, int dtype, const int32_t *symbols, const int32_t *ranges, int B, int S,
                      int T, int R, int C, int term, float *pxc, float *pyc, float *lse, cudaStream_t stream) {
  const int BTR = B * T * R;
  if (B > 65535) return FRN_EUNSUPPORTED;          // utterances ride on gridDim.y
  const dim3 grid((T * R + 7) / 8, B);
  if (dtype == FRN_F32)
    count_launch(), pruned_lse_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(logits), symbols, ranges, BTR,
                                                      T * R, S, R, C, term, pxc, pyc, lse);
  else if (dtype == FRN_BF16)
    count_launch(), pruned_lse_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(logits),
                                                              symbols, ranges, BTR, T * R, S, R, C, term, pxc,
                                                              pyc, lse);
  else return FRN_EINVAL;
  return check_launch();
}

int launch_skew_band(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary,
                     const DpGeom &g, const DpWorkspace &w, int R, int rnnt_type, float delay_penalty,
                     cudaStream_t stream) {
  // every arc of the lattice is dead except the band's
  const size_t cells = (size_t)g.B * g.Dn * g.P;
  count_launch(), fill_dead_arcs_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(w.XY, cells);
  BandParams bp{pxc, pyc, ranges, boundary, g.S, g.T, R, g.P, g.Dn, g.k, rnnt_type, delay_penalty};
  const int BTR = g.B * g.T * R;
  count_launch(), skew_band_kernel<<<(BTR + 255) / 256, 256, 0, stream>>>(bp, w.XY, BTR);
  return check_launch();
}

int launch_finalize_band(const int32_t *ranges, const int32_t *boundary, const DpGeom &g, const DpWorkspace &w,
                         int R, int rnnt_type, float *gxc, float *gyc, float *scores, cudaStream_t stream) {
  BandFinalizeParams fp{w.A, w.Bq, ranges, boundary, g.S, g.T, R, g.P, g.Dn, g.k, rnnt_type};
  const int n = gxc ? max(g.B * g.T * R, g.B) : g.B;
  count_launch(), finalize_band_kernel<<<(n + 255) / 256, 256, 0, stream>>>(fp, gxc, gyc, scores, g.B);
  return check_launch();
}

int launch_pruned_logits_grad(const void *logits, int dtype, const int32_t *symbols, const int32_t *ranges,
                              const float *lse, const float *gxc, const float *gyc, const float *scores_grad,
                              int B, int S, int T, int R, int C, int term, void *dlogits, cudaStream_t stream) {
  const int BTR = B * T * R;
  const int grid = (BTR + 7) / 8;
  if (dtype == FRN_F32)
    count_launch(), pruned_logits_grad_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(logits), symbols,
                                                              ranges, lse, gxc, gyc, scores_grad, BTR, T * R, S,
                                                              C, term, static_cast<float *>(dlogits));
  else if (dtype == FRN_BF16)
    count_launch(), pruned_logits_grad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(logits), symbols, ranges, lse, gxc, gyc, scores_grad, BTR, T * R, S,
        C, term, static_cast<__nv_bfloat16 *>(dlogits));
  else return FRN_EINVAL;
  return check_launch();
}

int launch_dense_to_band(const float *dpx, const float *dpy, const int32_t *ranges, const int32_t *boundary, int B,
                         int S, int T, int T1, int R, int rnnt_type, float *gxc, float *gyc, cudaStream_t stream) {
  const int n = B * T * R;
  count_launch(), dense_to_band_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dpx, dpy, ranges, boundary, B, S, T, T1, R,
                                                                         rnnt_type, gxc, gyc);
  return check_launch();
}

int launch_band_to_dense(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary,
                         int B, int S, int T, int T1, int R, int rnnt_type, float *px, float *py,
                         cudaStream_t stream) {
  const size_t n = (size_t)B * S * T1 + (size_t)B * (S + 1) * T;
  count_launch(), band_to_dense_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pxc, pyc, ranges, boundary, B, S, T, T1,
                                                                        R, rnnt_type, px, py);
  return check_launch();
}

}  // namespace frn
