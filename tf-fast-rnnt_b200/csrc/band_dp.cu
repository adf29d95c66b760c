// Lattice recursion restricted to the pruning band (A8), s_range <= 8.
//
// After pruning, frame t only has the R = s_range lattice rows
// ranges[b,t,0] .. ranges[b,t,0]+R-1 alive (rnnt_loss.py:968-1013 builds a dense
// [B,S,T+1] lattice that is -inf everywhere else and the reference then runs its
// dense kernels over it).  Here the recursion is expressed on the band itself as a
// product of R x R transfer matrices in the (logadd, +) semiring, one per frame:
//
//      alpha_{t+1} = M_t (x) alpha_t ,     beta_t = N_t (x) beta_{t+1}
//
// and evaluated in three short phases instead of S+T dependent steps:
//   1. every chunk of L = 16 frames propagates the R unit vectors -> chunk matrix
//      (R threads per chunk, 16 dependent frame steps);
//   2. one warp walks the chunk matrices (T/16 matrix-vector products) and leaves
//      the state at every chunk boundary;
//   3. every chunk replays its 16 frames from its boundary state and writes the
//      per-frame states.
// Dependent work: ~2*16 frame steps + T/16 matrix-vector products (~12k cycles at
// T=500) versus 600 wavefront steps.  Forward and backward directions run in
// different CTAs.  Numerics: every state vector is (exact integer offset, float32
// residuals), re-centred after each step, as in the dense chain kernel.
#include "common.cuh"
#include "launchers.h"

namespace frn {

constexpr int kBandR = 8;        // maximum band width handled here
constexpr int kBandChunk = 16;   // frames per chunk
constexpr int kBandThreads = 256;

struct BandDpParams {
  const float *pxc, *pyc;      // [B][T][R] natural-log band log-probs
  const int32_t *ranges;       // [B][T][R]
  const int32_t *boundary;     // [B][4]
  float *va, *ub;              // [B][T+1][kBandR] forward / backward residual states (log2 domain)
  float *oa, *ob;              // [B][T+1]          their integer offsets
  int S, T, R, modified, rnnt_type;
  float delay_penalty;
};

__device__ __forceinline__ int band_row_of(int r0, int i, int S1) {
  int s = (r0 + i) % S1;
  return s < 0 ? s + S1 : s;
}

// out[i] = (0 <= i + d < R) ? v[i + d] : kNeg      (d may be negative); entries >= R stay kNeg
template <int R>
__device__ __forceinline__ void shift_vec(const float (&v)[kBandR], int d, float (&out)[kBandR]) {
#pragma unroll
  for (int i = 0; i < kBandR; ++i) out[i] = kNeg;
  if (d == 0) {
#pragma unroll
    for (int i = 0; i < R; ++i) out[i] = v[i];
  } else {
#pragma unroll
    for (int dd = -(R - 1); dd < R; ++dd) {
      if (dd != 0 && d == dd) {
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (i + dd >= 0 && i + dd < R) out[i] = v[i + dd];
      }
    }
  }
}

// exact re-centring: move rint(max) into the offset
template <int R>
__device__ __forceinline__ void recentre(float (&v)[kBandR], float &off) {
  float m = v[0];
#pragma unroll
  for (int i = 1; i < R; ++i) m = fmaxf(m, v[i]);
  if (m > kNegThresh) {
    const float k = rintf(m);
    off += k;
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = (v[i] > kNegThresh) ? v[i] - k : kNeg;
  }
}

// log2( sum_j 2^z_j ) over n <= R finite terms (kNeg stands for -inf)
template <int N>
__device__ __forceinline__ float logsum2(const float (&z)[kBandR]) {
  float m = z[0];
#pragma unroll
  for (int j = 1; j < N; ++j) m = fmaxf(m, z[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) s += ex2_approx(z[j] - m);
  return m + lg2_approx(s);
}

// Per-utterance tables in shared memory (log2 domain, masks applied):
//   PX[t][i], PY[t][i] for t < Tb (arcs leaving column t), D[t] = r0[t+1] - r0[t]
//   (0 for the last transition: column Tb keeps the band of column Tb-1).
struct BandTables {
  const float *PX, *PY;
  const int *D;
  int Tb, modified;
};

// forward transition: state of column t -> state of column t+1
template <int R>
__device__ __forceinline__ void fwd_step(const BandTables &tb, int t, float (&v)[kBandR]) {
  const int d = tb.D[t];
  const float *py = tb.PY + t * kBandR, *px = tb.PX + t * kBandR;
  float a[kBandR], sh[kBandR];
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = v[i] + py[i];               // blank arcs (s,t)->(s,t+1)
  shift_vec<R>(a, d, sh);
  if (tb.modified) {
    // symbol arcs (s,t)->(s+1,t+1): destination index i' = i + 1 - d
    float bsrc[kBandR], b[kBandR];
#pragma unroll
    for (int i = 0; i < R; ++i) bsrc[i] = v[i] + px[i];
    shift_vec<R>(bsrc, d - 1, b);
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = logadd2(sh[i], b[i]);
  } else if (t + 1 < tb.Tb) {
    // symbol arcs stay in the column: closure along s with the px of column t+1, written
    // as R independent log-sum-exps  v[i'] = logsum_{j<=i'} ( sh[j] + px[j] + .. + px[i'-1] )
    const float *pxn = tb.PX + (t + 1) * kBandR;
    float pn[kBandR];
#pragma unroll
    for (int i = 0; i < R; ++i) pn[i] = pxn[i];
    float out[kBandR];
    out[0] = sh[0];
#pragma unroll
    for (int ip = 1; ip < R; ++ip) {
      float z[kBandR];
      float acc = 0.f;
      z[ip] = sh[ip];
#pragma unroll
      for (int j = ip - 1; j >= 0; --j) {
        acc += pn[j];
        z[j] = sh[j] + acc;
      }
      // terms z[0..ip]
      float m = z[0];
#pragma unroll
      for (int j = 1; j <= ip; ++j) m = fmaxf(m, z[j]);
      float ssum = 0.f;
#pragma unroll
      for (int j = 0; j <= ip; ++j) ssum += ex2_approx(z[j] - m);
      out[ip] = m + lg2_approx(ssum);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = out[i];
  } else {
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = sh[i];
  }
}

// backward transition: state of column t+1 -> state of column t
template <int R>
__device__ __forceinline__ void bwd_step(const BandTables &tb, int t, float (&u)[kBandR]) {
  const int d = tb.D[t];
  const float *py = tb.PY + t * kBandR, *px = tb.PX + t * kBandR;
  float sh[kBandR], w[kBandR];
  shift_vec<R>(u, -d, sh);                                       // u_{t+1}[i - d]
#pragma unroll
  for (int i = 0; i < R; ++i) w[i] = py[i] + sh[i];
  if (tb.modified) {
    float sh1[kBandR];
    shift_vec<R>(u, 1 - d, sh1);                                 // u_{t+1}[i + 1 - d]
#pragma unroll
    for (int i = 0; i < R; ++i) u[i] = logadd2(w[i], px[i] + sh1[i]);
  } else {
    // u_t[i] = logsum_{i'>=i} ( px[i] + .. + px[i'-1] + w[i'] ), R independent log-sum-exps
    float pn[kBandR];
#pragma unroll
    for (int i = 0; i < R; ++i) pn[i] = px[i];
    float out[kBandR];
    out[R - 1] = w[R - 1];
#pragma unroll
    for (int i = 0; i < R - 1; ++i) {
      float acc = 0.f;
      float m = w[i];
      float z[kBandR];
      z[i] = w[i];
#pragma unroll
      for (int ip = i + 1; ip < R; ++ip) {
        acc += pn[ip - 1];
        z[ip] = w[ip] + acc;
        m = fmaxf(m, z[ip]);
      }
      float ssum = 0.f;
#pragma unroll
      for (int ip = i; ip < R; ++ip) ssum += ex2_approx(z[ip] - m);
      out[i] = m + lg2_approx(ssum);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) u[i] = out[i];
  }
}

template <int R>
__global__ void __launch_bounds__(kBandThreads, 1) band_dp_kernel(BandDpParams p) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int T = p.T, S1 = p.S + 1;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, s_end = bd.z, t_end = bd.w;
  float *out_v = (dir ? p.ub : p.va) + (size_t)b * (T + 1) * kBandR;
  float *out_o = (dir ? p.ob : p.oa) + (size_t)b * (T + 1);
  const int Tb = t_end - t_begin;
  if (s_end < s_begin || Tb < 0 || s_begin < 0 || t_begin < 0 || s_end > p.S || t_end > T) return;

  const int nC = (Tb + kBandChunk - 1) / kBandChunk;             // chunks of transitions
  float *PX = reinterpret_cast<float *>(bsm);                    // [Tb+1][8]
  float *PY = PX + (size_t)(Tb + 1) * kBandR;                    // [Tb+1][8]
  int *D = reinterpret_cast<int *>(PY + (size_t)(Tb + 1) * kBandR);   // [Tb+1]
  int *R0 = D + (Tb + 1);                                        // [Tb+1]
  float *Pm = reinterpret_cast<float *>(R0 + (Tb + 1));          // [nC][8][8] chunk matrices (basis j -> row j)
  float *Poff = Pm + (size_t)nC * 64;                            // [nC][8]
  float *VB = Poff + (size_t)nC * kBandR;                        // [nC+1][8] boundary states
  float *VBo = VB + (size_t)(nC + 1) * kBandR;                   // [nC+1]

  // ---- phase 0: tables ----
  const float *pxc = p.pxc + (size_t)b * T * R, *pyc = p.pyc + (size_t)b * T * R;
  const int32_t *rg = p.ranges + (size_t)b * T * R;
  for (int t = tid; t <= Tb; t += kBandThreads) {
    const int ta = min(t_begin + (t < Tb ? t : max(Tb - 1, 0)), T - 1);   // column Tb keeps the band of Tb-1
    R0[t] = rg[(size_t)ta * R];
  }
  __syncthreads();
  for (int t = tid; t <= Tb; t += kBandThreads) D[t] = (t + 1 < Tb) ? R0[t + 1] - R0[t] : 0;
  for (int idx = tid; idx < (Tb + 1) * kBandR; idx += kBandThreads) {
    const int t = idx / kBandR, i = idx - t * kBandR;
    float vx = kNeg, vy = kNeg;
    if (t < Tb && i < R) {
      const int ta = t_begin + t;
      const int r0 = R0[t];
      const int s = band_row_of(r0, i, S1);
      if (s >= s_begin && s <= s_end) {
        vy = fmaxf(pyc[(size_t)ta * R + i] * kLog2e, kNeg);
        const bool next_in_band = (i + 1 < R) && band_row_of(r0, i + 1, S1) == s + 1;
        if (s < p.S && s < s_end && (p.modified || next_in_band)) {
          float v = pxc[(size_t)ta * R + i];
          if (p.rnnt_type == FRN_CONSTRAINED) v += next_in_band ? pyc[(size_t)ta * R + i + 1] : -INFINITY;
          if (p.delay_penalty != 0.f) v += delay_penalty_value(t_end, ta, p.delay_penalty);
          vx = fmaxf(v * kLog2e, kNeg);
        }
      }
    }
    PX[idx] = vx;
    PY[idx] = vy;
  }
  __syncthreads();
  BandTables tb{PX, PY, D, Tb, p.modified};

  // chunk c (in processing order of this direction) covers transitions
  //   forward : t in [c*L, min((c+1)*L, Tb))         column c*L   -> column min((c+1)L, Tb)
  //   backward: t in [max(Tb-(c+1)L,0), Tb - c*L)    column Tb-cL -> column max(Tb-(c+1)L, 0)
  // ---- phase 1: chunk matrices by propagating the unit vectors ----
  for (int w = tid; w < nC * R; w += kBandThreads) {
    const int c = w / R, j = w - c * R;
    float v[kBandR];
#pragma unroll
    for (int i = 0; i < kBandR; ++i) v[i] = (i == j) ? 0.f : kNeg;
    float off = 0.f;
    if (!dir) {
      const int t_lo = c * kBandChunk, t_hi = min(t_lo + kBandChunk, Tb);
      for (int t = t_lo; t < t_hi; ++t) { fwd_step<R>(tb, t, v); recentre<R>(v, off); }
    } else {
      const int t_hi = Tb - c * kBandChunk, t_lo = max(t_hi - kBandChunk, 0);
      for (int t = t_hi - 1; t >= t_lo; --t) { bwd_step<R>(tb, t, v); recentre<R>(v, off); }
    }
#pragma unroll
    for (int i = 0; i < kBandR; ++i) Pm[(size_t)c * 64 + j * kBandR + i] = v[i];
    Poff[c * kBandR + j] = off;
  }
  __syncthreads();

  // ---- phase 2: boundary states, one warp, lane i' <-> component ----
  if (tid < 32) {
    // initial state
    float x = kNeg;      // component `lane` of the current boundary state
    float off = 0.f;
    if (!dir) {
      const int i0 = s_begin - R0[0];
      float v0[kBandR];
#pragma unroll
      for (int i = 0; i < kBandR; ++i) v0[i] = (i == i0 && i0 >= 0 && i0 < R) ? 0.f : kNeg;
      if (!p.modified && Tb > 0) {               // symbols emitted on the first frame: closure in column 0
#pragma unroll
        for (int i = 1; i < kBandR; ++i) v0[i] = logadd2(v0[i - 1] + PX[i - 1], v0[i]);
      }
#pragma unroll
      for (int i = 0; i < kBandR; ++i) if (lane == i) x = v0[i];
    } else {
      const int iE = s_end - R0[Tb];
      x = (lane == iE && iE >= 0 && iE < R) ? 0.f : kNeg;
    }
    if (lane < kBandR) VB[lane] = x;
    if (lane == 0) VBo[0] = 0.f;
    for (int c = 0; c < nC; ++c) {
      // y[i'] = logsum_j ( x[j] + Pm[c][j][i'] + Poff[c][j] )
      float z[kBandR], m = kNeg;
#pragma unroll
      for (int j = 0; j < kBandR; ++j) {
        const float xj = __shfl_sync(0xffffffffu, x, j);
        z[j] = (j < R && lane < kBandR) ? xj + (Pm[(size_t)c * 64 + j * kBandR + lane] + Poff[c * kBandR + j]) : kNeg;
        m = fmaxf(m, z[j]);
      }
      float y = kNeg;
      if (m > kNegThresh) {
        float ssum = 0.f;
#pragma unroll
        for (int j = 0; j < kBandR; ++j) ssum += ex2_approx(z[j] - m);
        y = m + lg2_approx(ssum);
      }
      // exact re-centring by rint(max over components)
      float mm = y;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
      if (mm > kNegThresh) {
        const float k = rintf(mm);
        off += k;
        y = (y > kNegThresh) ? y - k : kNeg;
      }
      x = y;
      if (lane < kBandR) VB[(c + 1) * kBandR + lane] = x;
      if (lane == 0) VBo[c + 1] = off;
    }
  }
  __syncthreads();

  // ---- phase 3: replay every chunk from its boundary state, write per-column states ----
  for (int c = tid; c < nC + 1; c += kBandThreads) {
    float v[kBandR];
#pragma unroll
    for (int i = 0; i < kBandR; ++i) v[i] = VB[c * kBandR + i];
    float off = VBo[c];
    if (!dir) {
      const int t_lo = min(c * kBandChunk, Tb);    // column of this boundary state
      const int t_hi = min(t_lo + kBandChunk, Tb);
      int col = t_lo;
      *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR + 4) = make_float4(v[4], v[5], v[6], v[7]);
      out_o[col] = off;
      if (c == nC) continue;
      for (int t = t_lo; t < t_hi - 1; ++t) {      // the last column of the chunk belongs to the next boundary
        fwd_step<R>(tb, t, v); recentre<R>(v, off);
        col = t + 1;
        *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR + 4) = make_float4(v[4], v[5], v[6], v[7]);
        out_o[col] = off;
      }
    } else {
      const int t_hi = max(Tb - c * kBandChunk, 0);   // column of this boundary state
      const int t_lo = max(t_hi - kBandChunk, 0);
      int col = t_hi;
      *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR + 4) = make_float4(v[4], v[5], v[6], v[7]);
      out_o[col] = off;
      if (c == nC) continue;
      for (int t = t_hi - 1; t > t_lo; --t) {
        bwd_step<R>(tb, t, v); recentre<R>(v, off);
        col = t;
        *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(out_v + (size_t)col * kBandR + 4) = make_float4(v[4], v[5], v[6], v[7]);
        out_o[col] = off;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// scores + compact occupation counts from the band states.  Thread per (b,t,i).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) band_finalize_kernel(BandDpParams p, float *gxc, float *gyc, float *scores,
                                                            int B) {
  const int T = p.T, R = p.R, S1 = p.S + 1, TR = T * R;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  auto ok_bd = [&](const int4 &bd) {
    return bd.z >= bd.x && bd.w >= bd.y && bd.x >= 0 && bd.y >= 0 && bd.z <= p.S && bd.w <= T;
  };
  auto total_of = [&](int b, const int4 &bd, float &tr, float &to) -> bool {
    const int Tb = bd.w - bd.y;
    const int r0e = p.ranges[(size_t)(b * T + min(bd.y + max(Tb - 1, 0), T - 1)) * R];
    const int iE = bd.z - r0e;
    if (iE < 0 || iE >= R) return false;
    tr = p.va[((size_t)b * (T + 1) + Tb) * kBandR + iE];
    to = p.oa[(size_t)b * (T + 1) + Tb];
    return tr > kNegThresh;
  };
  if (idx < B && scores) {
    const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * idx);
    float v = 0.f;
    if (ok_bd(bd)) {
      float tr, to;
      v = total_of(idx, bd, tr, to) ? (float)(((double)tr + (double)to) * 0.6931471805599453) : -INFINITY;
    }
    scores[idx] = v;
  }
  if (idx >= B * TR || gxc == nullptr) return;
  const int b = idx / TR, rem = idx - b * TR, ta = rem / R, i = rem - ta * R;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  float vx = 0.f, vy = 0.f;
  float tr, to;
  if (ok_bd(bd) && ta >= bd.y && ta < bd.w && total_of(b, bd, tr, to)) {
    const int t = ta - bd.y, Tb = bd.w - bd.y;
    const int r0 = p.ranges[(size_t)(b * T + ta) * R];
    const int d = (t + 1 < Tb) ? p.ranges[(size_t)(b * T + ta + 1) * R] - r0 : 0;
    const size_t base = (size_t)b * (T + 1);
    const float *va = p.va + (base + t) * kBandR, *un = p.ub + (base + t + 1) * kBandR, *uc = p.ub + (base + t) * kBandR;
    const float oa = p.oa[base + t], obn = p.ob[base + t + 1], obc = p.ob[base + t];
    auto arc = [&](int ii, bool px_arc) -> float {
      const int s = band_row_of(r0, ii, S1);
      if (s < bd.x || s > bd.z) return 0.f;
      const float a = va[ii];
      if (!(a > kNegThresh)) return 0.f;
      float score, head, ohead;
      if (!px_arc) {
        const int ih = ii - d;
        if (ih < 0 || ih >= R) return 0.f;
        score = p.pyc[(size_t)(b * T + ta) * R + ii];
        head = un[ih]; ohead = obn;
      } else {
        const bool next_in_band = (ii + 1 < R) && band_row_of(r0, ii + 1, S1) == s + 1;
        if (!(s < p.S && s < bd.z && (p.modified || next_in_band))) return 0.f;
        score = p.pxc[(size_t)(b * T + ta) * R + ii];
        if (p.rnnt_type == FRN_CONSTRAINED) score += next_in_band ? p.pyc[(size_t)(b * T + ta) * R + ii + 1] : -INFINITY;
        if (p.delay_penalty != 0.f) score += delay_penalty_value(bd.w, ta, p.delay_penalty);
        if (p.modified) {
          const int ih = ii + 1 - d;
          if (ih < 0 || ih >= R) return 0.f;
          head = un[ih]; ohead = obn;
        } else {
          head = uc[ii + 1]; ohead = obc;
        }
      }
      if (!(head > kNegThresh) || !(score > -INFINITY)) return 0.f;
      const float e = ((a - tr) + head + score * kLog2e) + ((oa + ohead) - to);
      return ex2_approx(e);
    };
    vx = arc(i, true);
    vy = arc(i, false);
    if (p.rnnt_type == FRN_CONSTRAINED && i >= 1 && band_row_of(r0, i - 1, S1) == band_row_of(r0, i, S1) - 1)
      vy += arc(i - 1, true);      // the px arc of row s-1 borrowed py[s,t]
  }
  gxc[idx] = vx;
  gyc[idx] = vy;
}

// ---------------------------------------------------------------------------
size_t band_dp_workspace_bytes(int B, int T) {
  return 2 * round_up_sz((size_t)B * (T + 1) * kBandR * sizeof(float), 256) +
         2 * round_up_sz((size_t)B * (T + 1) * sizeof(float), 256);
}

bool band_dp_supported(int S, int T, int R) {
  if (R > kBandR) return false;
  const int nC = (T + kBandChunk - 1) / kBandChunk;
  const size_t smem = (size_t)(T + 1) * (2 * kBandR * 4 + 8) + (size_t)nC * (64 + kBandR) * 4 + (size_t)(nC + 1) * (kBandR + 1) * 4 + 64;
  return smem <= 200 * 1024;
}

int launch_band_dp(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary, int B, int S,
                   int T, int R, int rnnt_type, float delay_penalty, bool want_grad, void *workspace, float *gxc,
                   float *gyc, float *scores, cudaStream_t stream) {
  char *w = static_cast<char *>(workspace);
  const size_t nv = round_up_sz((size_t)B * (T + 1) * kBandR * sizeof(float), 256);
  const size_t no = round_up_sz((size_t)B * (T + 1) * sizeof(float), 256);
  BandDpParams p;
  p.pxc = pxc; p.pyc = pyc; p.ranges = ranges; p.boundary = boundary;
  p.va = reinterpret_cast<float *>(w); p.ub = reinterpret_cast<float *>(w + nv);
  p.oa = reinterpret_cast<float *>(w + 2 * nv); p.ob = reinterpret_cast<float *>(w + 2 * nv + no);
  p.S = S; p.T = T; p.R = R; p.modified = (rnnt_type != FRN_REGULAR); p.rnnt_type = rnnt_type;
  p.delay_penalty = delay_penalty;
  const int nC = (T + kBandChunk - 1) / kBandChunk;
  const size_t smem = (size_t)(T + 1) * (2 * kBandR * 4 + 8) + (size_t)nC * (64 + kBandR) * 4 +
                      (size_t)(nC + 1) * (kBandR + 1) * 4 + 64;
  cudaError_t e = cudaSuccess;
#define FRN_BAND(R_)                                                                                       \
  case R_:                                                                                                 \
    e = cudaFuncSetAttribute(band_dp_kernel<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) return note_cuda_error(e);                                                       \
    band_dp_kernel<R_><<<dim3(B, want_grad ? 2 : 1), kBandThreads, smem, stream>>>(p);                     \
    break;
  switch (R) {
    FRN_BAND(1) FRN_BAND(2) FRN_BAND(3) FRN_BAND(4) FRN_BAND(5) FRN_BAND(6) FRN_BAND(7) FRN_BAND(8)
    default: return FRN_EUNSUPPORTED;
  }
#undef FRN_BAND
  int rc = check_launch();
  if (rc) return rc;
  const int n = want_grad ? max(B * T * R, B) : B;
  band_finalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p, want_grad ? gxc : nullptr, want_grad ? gyc : nullptr,
                                                            scores, B);
  return check_launch();
}

}  // namespace frn
