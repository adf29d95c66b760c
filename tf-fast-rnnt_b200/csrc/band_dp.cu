// Lattice recursion restricted to the pruning band (A8), s_range <= 8.
//
// After pruning, frame t only has the R = s_range lattice rows
// ranges[b,t,0] .. ranges[b,t,0]+R-1 alive (rnnt_loss.py:968-1013 builds a dense
// [B,S,T+1] lattice that is -inf everywhere else and the reference then runs its
// dense kernels over it).  Here the recursion is expressed on the band itself as a
// product of 8 x 8 transfer matrices, one per frame,
//
//      alpha_{t+1} = M_t alpha_t ,     beta_t = N_t beta_{t+1}
//
// in the LINEAR domain: a band column is 8 float64 mantissas sharing one exact
// int32 frame (value = m * 2^frame, re-normalised by an exact power of two after
// every frame), so a frame costs a handful of FMAs and no transcendental.
//
// Slots.  Lattice row s always lives in slot s & 7.  As the band moves up with t
// a row keeps its slot, rows that leave the band are killed by a zero arc and
// rows that enter start at zero, so nothing is ever shifted: all band geometry
// (ranges, boundary, rnnt_type, delay penalty) is folded into two per-frame arc
// tables in shared memory, PY[t][k] (blank arc of the row in slot k) and
// PX[t][k] (symbol arc), built once per utterance.  The within-frame closure of
// the regular recursion (several symbols on one frame) is a cyclic first-order
// recurrence over the slots with at least one zero link; it is solved exactly by
// three doubling steps.
//
// Evaluation in three short phases instead of S+T dependent steps:
//   1. every chunk of L frames propagates the 8 unit vectors -> chunk matrix
//      (8 threads per chunk, L dependent frame steps);
//   2. eight lanes walk the chunk matrices (T/L matrix-vector products) and leave
//      the state at every chunk boundary;
//   3. every column in parallel: its state is the image (recorded in phase 1) of its
//      chunk's boundary state, one 8x8 matrix-vector product; written (mantissas +
//      frame) for the finalize kernel.
// Forward and backward directions run in different CTAs.
//
// Limits: band entries whose row index wraps around S+1 (only possible with
// ranges the reference never produces) are treated as outside the band; entries
// of one column that are more than 2^1000 apart flush to zero.
#include "common.cuh"
#include "launchers.h"
#ifdef FRN_BAND_TIMING   // diagnostic build: block (0,0) prints the cycle count of every phase
#include <cstdio>
#define BAND_T(i) do { if (tid == 0 && blockIdx.x == 0) tclk[i] = clock64(); } while (0)
#else
#define BAND_T(i) do { } while (0)
#endif

namespace frn {

constexpr int kBandR = 8;          // maximum band width handled here = number of slots
constexpr int kBandThreads = 512;
#ifndef FRN_BAND_MIN_CHUNK
#define FRN_BAND_MIN_CHUNK 16
#endif
constexpr int kBandMinChunk = FRN_BAND_MIN_CHUNK;   // frames per chunk (doubled until the tables fit shared memory)
constexpr int kDeadFrame = -(1 << 29);

struct BandDpParams {
  const float *pxc, *pyc;      // [B][T][R] natural-log band log-probs
  const int32_t *ranges;       // [B][T][R]
  const int32_t *boundary;     // [B][4]
  double *va, *ub;             // [B][T+1][8] forward / backward states by slot: mantissas
  int *oa, *ob;                // [B][T+1]    their frames (kDeadFrame: all-zero state)
  uint32_t *img;               // [B][2][T+1][8 unit vectors][8] image of chunk-start unit vector j at this column:
                               // the HIGH WORD of the float64 entry, rounded (sign, 11-bit exponent, 20-bit
                               // mantissa).  The full exponent range matters - with a delay penalty the entries
                               // of one image are e^(penalty * symbols apart) from each other, far beyond
                               // float32's 2^126 (float32 images flushed them: wrong occupation counts) - the
                               // mantissa does not: 2^-21 per entry, float32 storage cost, float64 range.
  int *img_frame;              // [B][2][T+1][8]  its frame (kDeadFrame: zero image)
  int S, T, R, L, modified, rnnt_type;
  float delay_penalty;
};

// 2^e as a double for e in [-1022, 1023]; +0 below
__device__ __forceinline__ double pow2d(int e) { return __hiloint2double((max(e, -1023) + 1023) << 20, 0); }

// probability of an arc from its natural-log score: mantissa by exp2f (float accuracy, as in the
// dense chain), exponent exact.  Scores below -1000 (log2) count as -inf.
__device__ __forceinline__ double arc_prob(float v) {
  const float x = v * kLog2e;
  if (!(x > -1000.f)) return 0.0;
  const float e = floorf(x);
  return (double)exp2f(x - e) * pow2d((int)e);
}

// v *= 2^-k with k = exponent of the largest entry; returns k (kDeadFrame for the zero vector)
__device__ __forceinline__ int normalise8(double (&v)[8]) {
  int hi = __double2hiint(v[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k) hi = max(hi, __double2hiint(v[k]));   // entries are >= 0: integer order = magnitude order
  if ((hi >> 20) == 0) {                                           // zero (or denormal): dead
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.0;
    return kDeadFrame;
  }
  const int e = (hi >> 20) - 1023;
  const double sc = pow2d(-e);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] *= sc;
  return e;
}

// x[k] <- sum_{j>=0} x[k -/+ j] * prod of the j links leading to k : cyclic first-order recurrence
// x[k] = b[k] + a[k] * x[k - DIRN] with at least one zero link, by doubling.  a[k] is the link INTO k.
template <int DIRN>
__device__ __forceinline__ void closure8(double (&x)[8], const double (&a1)[8]) {
  double a2[8], a4[8], y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a2[k] = a1[k] * a1[(k - DIRN) & 7];
#pragma unroll
  for (int k = 0; k < 8; ++k) y[k] = fma(a1[k], x[(k - DIRN) & 7], x[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k) a4[k] = a2[k] * a2[(k - 2 * DIRN) & 7];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = fma(a2[k], y[(k - 2 * DIRN) & 7], y[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k) y[k] = fma(a4[k], x[(k - 4 * DIRN) & 7], x[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = y[k];
}

struct BandTables {
  const double *PX, *PY;   // [Tb+1][8]
  int modified;
};

__device__ __forceinline__ void load8(const double *src, double (&d)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const double2 t = *reinterpret_cast<const double2 *>(src + k);
    d[k] = t.x; d[k + 1] = t.y;
  }
}

// forward transition: closed state of column t -> closed state of column t+1
__device__ __forceinline__ void fwd_step(const BandTables &tb, int t, double (&v)[8]) {
  double py[8], px[8];
  load8(tb.PY + t * 8, py);
  if (tb.modified) {
    load8(tb.PX + t * 8, px);
    double n[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) n[k] = fma(v[(k - 1) & 7], px[(k - 1) & 7], v[k] * py[k]);   // (s,t)->(s+1,t+1) and blank
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = n[k];
  } else {
    load8(tb.PX + (t + 1) * 8, px);
    double a1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] *= py[k]; a1[k] = px[(k - 1) & 7]; }   // link into slot k: symbol arc of slot k-1
    closure8<1>(v, a1);
  }
}

// backward transition: state of column t+1 -> state of column t
__device__ __forceinline__ void bwd_step(const BandTables &tb, int t, double (&u)[8]) {
  double py[8], px[8];
  load8(tb.PY + t * 8, py);
  load8(tb.PX + t * 8, px);
  if (tb.modified) {
    double n[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) n[k] = fma(px[k], u[(k + 1) & 7], py[k] * u[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = n[k];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] *= py[k];
    closure8<-1>(u, px);                                                       // link into slot k from slot k+1
  }
}

__device__ __forceinline__ void store_state(const double (&v)[8], int frame, double *out_v, int *out_o) {
#pragma unroll
  for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2 *>(out_v + k) = make_double2(v[k], v[k + 1]);
  *out_o = frame;
}

__host__ __device__ inline size_t band_smem_bytes(int T, int L) {
  const size_t nC = (size_t)(T + L - 1) / L;
  return (size_t)(T + 1) * 16 * sizeof(double)        // PX, PY
         + (size_t)(T + 2) * sizeof(int)               // R0
         + nC * 64 * sizeof(double)                    // chunk matrices
         + nC * 8 * sizeof(int) + nC * sizeof(int)     // per-vector frames, chunk frames
         + (nC + 1) * 8 * sizeof(double) + (nC + 1) * sizeof(int) + 64;   // boundary states
}

__global__ void __launch_bounds__(kBandThreads, 1) band_dp_kernel(BandDpParams p) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
  const int T = p.T, R = p.R, L = p.L;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, s_end = bd.z, t_end = bd.w;
  double *out_v = (dir ? p.ub : p.va) + (size_t)b * (T + 1) * kBandR;
  int *out_o = (dir ? p.ob : p.oa) + (size_t)b * (T + 1);
  const int Tb = t_end - t_begin;
  if (s_end < s_begin || Tb < 0 || s_begin < 0 || t_begin < 0 || s_end > p.S || t_end > T) return;

  const int nC = (Tb + L - 1) / L;                                   // chunks of transitions
  double *PX = reinterpret_cast<double *>(bsm);                      // [Tb+1][8]
  double *PY = PX + (size_t)(Tb + 1) * 8;                            // [Tb+1][8]
  double *Pm = PY + (size_t)(Tb + 1) * 8;                            // [nC][8][8] chunk matrices (basis j -> row j)
  double *VB = Pm + (size_t)nC * 64;                                 // [nC+1][8] boundary states
  int *R0 = reinterpret_cast<int *>(VB + (size_t)(nC + 1) * 8);      // [Tb+2]
  int *Fv = R0 + (Tb + 2);                                           // [nC][8] frame of each propagated unit vector
  int *Poff = Fv + (size_t)nC * 8;                                   // [nC]    common frame of a chunk matrix
  int *VBo = Poff + nC;                                              // [nC+1]

#ifdef FRN_BAND_TIMING
  long long tclk[6];
#endif
  BAND_T(0);
  // ---- phase 0: arc tables in slot order ----
  const float *__restrict__ pxc = p.pxc + (size_t)b * T * R;
  const float *__restrict__ pyc = p.pyc + (size_t)b * T * R;
  const int32_t *__restrict__ rg = p.ranges + (size_t)b * T * R;
  for (int t = tid; t <= Tb + 1; t += kBandThreads) {
    const int ta = min(t_begin + (t < Tb ? t : max(Tb - 1, 0)), T - 1);   // columns >= Tb keep the band of Tb-1
    R0[t] = rg[(size_t)ta * R];
  }
  __syncthreads();
  // (loads are unconditional on clamped indices so that the unrolled iterations issue them together)
#pragma unroll 4
  for (int idx = tid; idx < (Tb + 1) * 8; idx += kBandThreads) {
    const int t = idx >> 3, k = idx & 7;
    const int tt = min(t, max(Tb - 1, 0));
    const int r0 = R0[tt], r1 = R0[tt + 1];
    const int i = (k - r0) & 7, s = r0 + i;
    const int ic = min(i, R - 1);
    const int ta = min(t_begin + tt, T - 1);
    const float fy = pyc[(size_t)ta * R + ic];
    const float fx0 = pxc[(size_t)ta * R + ic];
    const float fy1 = pyc[(size_t)ta * R + min(ic + 1, R - 1)];
    double vx = 0.0, vy = 0.0;
    if (t < Tb && i < R && s >= s_begin && s <= s_end) {                // (s <= s_end <= S: no wrap-around)
      const bool next_in_band = (i + 1 < R);
      float fx = -INFINITY;
      if (s < p.S && s < s_end) {
        fx = fx0;
        if (p.rnnt_type == FRN_CONSTRAINED) fx += next_in_band ? fy1 : -INFINITY;
        if (p.delay_penalty != 0.f) fx += delay_penalty_value(t_end, ta, p.delay_penalty);
      }
      if ((unsigned)(s - r1) < (unsigned)R) vy = arc_prob(fy);          // same row still inside the band of column t+1
      if (p.modified) {
        if ((unsigned)(s + 1 - r1) < (unsigned)R) vx = arc_prob(fx);   // row s+1 inside the band of column t+1
      } else if (next_in_band) {
        vx = arc_prob(fx);
      }
    }
    PX[idx] = vx;
    PY[idx] = vy;
  }
  __syncthreads();
  BAND_T(1);
  const BandTables tb{PX, PY, p.modified};

  // chunk c (in processing order of this direction) covers transitions
  //   forward : t in [c*L, min((c+1)*L, Tb))         column c*L   -> column min((c+1)L, Tb)
  //   backward: t in [max(Tb-(c+1)L,0), Tb - c*L)    column Tb-cL -> column max(Tb-(c+1)L, 0)
  // ---- phase 1: chunk matrices by propagating the unit vectors ----
  {
    const int w = tid, c = w >> 3, j = w & 7;
    double v[8];
    int frame = kDeadFrame;
    if (w < nC * 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (k == j) ? 1.0 : 0.0;
      frame = 0;
      // every intermediate image is kept (rounded float64 high words + frame): the state of a column inside a
      // chunk is then one 8x8 matrix-vector product with the chunk's boundary state (phase 3)
      uint32_t *img = p.img + ((size_t)(b * 2 + dir) * (T + 1)) * 64 + j * 8;
      int *imf = p.img_frame + ((size_t)(b * 2 + dir) * (T + 1)) * 8 + j;
      auto record = [&](int col) {
        uint32_t *q = img + (size_t)col * 64;
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)      // entries are >= 0: round the magnitude to the nearest 2^-20 of its binade
          h[k] = (uint32_t)__double2hiint(v[k]) + ((uint32_t)__double2loint(v[k]) >> 31);
        *reinterpret_cast<uint4 *>(q) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(q + 4) = make_uint4(h[4], h[5], h[6], h[7]);
        imf[(size_t)col * 8] = frame;
      };
      if (!dir) {
        const int t_lo = c * L, t_hi = min(t_lo + L, Tb);
        for (int t = t_lo; t < t_hi; ++t) {
          fwd_step(tb, t, v);
          const int e = normalise8(v);
          frame = (e == kDeadFrame || frame <= kDeadFrame / 2) ? kDeadFrame : frame + e;
          record(t + 1);
        }
      } else {
        const int t_hi = Tb - c * L, t_lo = max(t_hi - L, 0);
        for (int t = t_hi - 1; t >= t_lo; --t) {
          bwd_step(tb, t, v);
          const int e = normalise8(v);
          frame = (e == kDeadFrame || frame <= kDeadFrame / 2) ? kDeadFrame : frame + e;
          record(t);
        }
      }
      Fv[w] = frame;
    }
    __syncthreads();
    if (w < nC * 8) {
      int F = Fv[c * 8];
#pragma unroll
      for (int q = 1; q < 8; ++q) F = max(F, Fv[c * 8 + q]);
      const double sc = (frame > kDeadFrame / 2) ? pow2d(frame - F) : 0.0;   // all eight images in one frame
#pragma unroll
      for (int k = 0; k < 8; ++k) Pm[(size_t)c * 64 + j * 8 + k] = v[k] * sc;
      if (j == 0) Poff[c] = F;
    }
  }
  __syncthreads();

  BAND_T(2);
  // ---- phase 2: boundary states; lanes 0..7 of warp 0 each own one slot of the state ----
  if (tid < 32) {
    const int lane = tid, k = lane & 7;
    double x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = 0.0;
    if (!dir) {
      const unsigned i0 = (unsigned)(s_begin - R0[0]);
      if (i0 < (unsigned)R) {
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (q == (s_begin & 7)) ? 1.0 : 0.0;
      }
      if (!p.modified && Tb > 0) {               // symbols emitted on the first frame: closure in column 0
        double px[8], a1[8];
        load8(PX, px);
#pragma unroll
        for (int q = 0; q < 8; ++q) a1[q] = px[(q - 1) & 7];
        closure8<1>(x, a1);
      }
    } else {
      const unsigned iE = (unsigned)(s_end - R0[Tb]);
      if (iE < (unsigned)R) {
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (q == (s_end & 7)) ? 1.0 : 0.0;
      }
    }
    int frame = normalise8(x);                   // every lane holds the whole (identical) state
    auto mine = [&]() {                          // x[k] without a dynamically indexed register array
      double r = x[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) r = (k == q) ? x[q] : r;
      return r;
    };
    if (lane < 8) VB[lane] = mine();
    if (lane == 0) VBo[0] = frame;
    for (int c = 0; c < nC; ++c) {
      // y[k] = sum_j x[j] * M[c][j][k] on lane k, two partial sums to halve the dependent chain
      const double *M = Pm + (size_t)c * 64 + k;
      double y0 = 0.0, y1 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        y0 = fma(x[j], M[j * 8], y0);
        y1 = fma(x[j + 1], M[(j + 1) * 8], y1);
      }
      const double yk = y0 + y1;
      // all-gather the eight slots (lanes 8..31 mirror lanes 0..7)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int hi = __shfl_sync(0xffffffffu, __double2hiint(yk), q);
        const int lo = __shfl_sync(0xffffffffu, __double2loint(yk), q);
        x[q] = __hiloint2double(hi, lo);
      }
      const int F = Poff[c];
      const int e = normalise8(x);
      frame = (e == kDeadFrame || F <= kDeadFrame / 2 || frame <= kDeadFrame / 2) ? kDeadFrame : frame + F + e;
      if (lane < 8) VB[(size_t)(c + 1) * 8 + lane] = mine();
      if (lane == 0) VBo[c + 1] = frame;
    }
  }
  __syncthreads();

  BAND_T(3);
  // ---- phase 3: every column in parallel: state = (recorded images of its chunk) x (boundary state) ----
  __threadfence_block();
  for (int t = tid; t <= Tb; t += kBandThreads) {
    // chunk whose boundary state feeds column t, and whether t is that boundary itself
    const int dist = dir ? Tb - t : t;                 // transitions between the direction's start column and t
    const int c = (dist == 0) ? 0 : (dist - 1) / L;    // column t is produced by a transition of chunk c ...
    const bool boundary_col = (dist % L == 0) || (dist == Tb);   // ... or is a stored boundary state
    double y[8];
    int frame;
    if (boundary_col) {
      const int cb = (dist == Tb) ? nC : dist / L;
      load8(VB + (size_t)cb * 8, y);
      frame = VBo[cb];
    } else {
      double x[8];
      load8(VB + (size_t)c * 8, x);
      const int fx = VBo[c];
      const uint32_t *img = p.img + ((size_t)(b * 2 + dir) * (T + 1) + t) * 64;
      const int *imf = p.img_frame + ((size_t)(b * 2 + dir) * (T + 1) + t) * 8;
      int fj[8], E = kDeadFrame;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        fj[j] = (x[j] > 0.0) ? imf[j] : kDeadFrame;
        E = max(E, fj[j]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = 0.0;
      if (E > kDeadFrame / 2 && fx > kDeadFrame / 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double xs = (fj[j] > kDeadFrame / 2) ? x[j] * pow2d(fj[j] - E) : 0.0;
          const uint4 lo = *reinterpret_cast<const uint4 *>(img + j * 8);
          const uint4 hi = *reinterpret_cast<const uint4 *>(img + j * 8 + 4);
          const uint32_t h[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = fma(xs, __hiloint2double((int)h[k], 0), y[k]);
        }
        const int e = normalise8(y);
        frame = (e == kDeadFrame) ? kDeadFrame : fx + E + e;
      } else {
        frame = kDeadFrame;
      }
    }
    store_state(y, frame, out_v + (size_t)t * kBandR, out_o + t);
  }
  BAND_T(4);
#ifdef FRN_BAND_TIMING
  if (tid == 0 && blockIdx.x == 0)
    printf("band_dp dir %d Tb %d: tables %lld, unit vectors %lld, boundary walk %lld, columns %lld cycles\n", dir, Tb,
           tclk[1] - tclk[0], tclk[2] - tclk[1], tclk[3] - tclk[2], tclk[4] - tclk[3]);
#endif
}

// ---------------------------------------------------------------------------
// scores + compact occupation counts from the band states.  Thread per (b,t,i).
// occupation of an arc = alpha(tail) * P(arc) * beta(head) / total, all in the linear domain.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) band_finalize_kernel(BandDpParams p, float *gxc, float *gyc, float *scores,
                                                            int B) {
  const int T = p.T, R = p.R, TR = T * R;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  auto ok_bd = [&](const int4 &bd) {
    return bd.z >= bd.x && bd.w >= bd.y && bd.x >= 0 && bd.y >= 0 && bd.z <= p.S && bd.w <= T;
  };
  auto total_of = [&](int b, const int4 &bd, double &tm, int &to) -> bool {
    const int Tb = bd.w - bd.y;
    const int r0e = p.ranges[(size_t)(b * T + min(bd.y + max(Tb - 1, 0), T - 1)) * R];
    if ((unsigned)(bd.z - r0e) >= (unsigned)R) return false;
    tm = p.va[((size_t)b * (T + 1) + Tb) * kBandR + (bd.z & 7)];
    to = p.oa[(size_t)b * (T + 1) + Tb];
    return tm > 0.0 && to > kDeadFrame / 2;
  };
  if (idx < B && scores) {
    const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * idx);
    float v = 0.f;
    if (ok_bd(bd)) {
      double tm;
      int to;
      v = -INFINITY;
      if (total_of(idx, bd, tm, to)) {
        const int e = ((__double2hiint(tm) >> 20) & 0x7ff) - 1023;
        v = (float)(((double)log2f((float)(tm * pow2d(-e))) + (double)(e + to)) * 0.6931471805599453);
      }
    }
    scores[idx] = v;
  }
  if (idx >= B * TR || gxc == nullptr) return;
  const int b = idx / TR, rem = idx - b * TR, ta = rem / R, i = rem - ta * R;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  float vx = 0.f, vy = 0.f;
  double tm;
  int to;
  if (ok_bd(bd) && ta >= bd.y && ta < bd.w && total_of(b, bd, tm, to)) {
    const int t = ta - bd.y, Tb = bd.w - bd.y;
    const int r0 = p.ranges[(size_t)(b * T + ta) * R];
    const int d = (t + 1 < Tb) ? p.ranges[(size_t)(b * T + ta + 1) * R] - r0 : 0;
    const size_t base = (size_t)b * (T + 1);
    const double *va = p.va + (base + t) * kBandR, *un = p.ub + (base + t + 1) * kBandR, *uc = p.ub + (base + t) * kBandR;
    const int oa = p.oa[base + t], obn = p.ob[base + t + 1], obc = p.ob[base + t];
    const double inv_tot = 1.0 / tm;
    auto arc = [&](int ii, bool px_arc) -> float {
      const int s = r0 + ii;                       // rows that would wrap around S+1 are outside the band
      if (s < bd.x || s > bd.z) return 0.f;
      const double a = va[s & 7];
      if (!(a > 0.0) || oa <= kDeadFrame / 2) return 0.f;
      float score;
      double head;
      int ohead;
      if (!px_arc) {
        if ((unsigned)(ii - d) >= (unsigned)R) return 0.f;
        score = p.pyc[(size_t)(b * T + ta) * R + ii];
        head = un[s & 7]; ohead = obn;
      } else {
        const bool next_in_band = (ii + 1 < R);
        if (!(s < p.S && s < bd.z && (p.modified || next_in_band))) return 0.f;
        score = p.pxc[(size_t)(b * T + ta) * R + ii];
        if (p.rnnt_type == FRN_CONSTRAINED) score += next_in_band ? p.pyc[(size_t)(b * T + ta) * R + ii + 1] : -INFINITY;
        if (p.delay_penalty != 0.f) score += delay_penalty_value(bd.w, ta, p.delay_penalty);
        if (p.modified) {
          if ((unsigned)(ii + 1 - d) >= (unsigned)R) return 0.f;
          head = un[(s + 1) & 7]; ohead = obn;
        } else {
          head = uc[(s + 1) & 7]; ohead = obc;
        }
      }
      if (!(head > 0.0) || ohead <= kDeadFrame / 2) return 0.f;
      // The mantissas are <= 2 but may be tiny relative to their column's frame (an entry far
      // from the column's mass), so the frame difference can be large and positive; the true
      // occupation is <= 1, so the product stays finite, and it flushes to 0 when negligible.
      return (float)((((a * inv_tot) * head) * arc_prob(score)) * pow2d(min(oa + ohead - to, 1000)));
    };
    vx = arc(i, true);
    vy = arc(i, false);
    if (p.rnnt_type == FRN_CONSTRAINED && i >= 1) vy += arc(i - 1, true);      // the px arc of row s-1 borrowed py[s,t]
  }
  gxc[idx] = vx;
  gyc[idx] = vy;
}

// ---------------------------------------------------------------------------
size_t band_dp_workspace_bytes(int B, int T) {
  return 2 * round_up_sz((size_t)B * (T + 1) * kBandR * sizeof(double), 256) +
         2 * round_up_sz((size_t)B * (T + 1) * sizeof(int), 256) +
         round_up_sz((size_t)B * 2 * (T + 1) * 64 * sizeof(uint32_t), 256) +
         round_up_sz((size_t)B * 2 * (T + 1) * 8 * sizeof(int), 256);
}

static int band_chunk_len(int T) {
  for (int L = kBandMinChunk; L <= 256; L <<= 1)
    if ((T + L - 1) / L * 8 <= kBandThreads && band_smem_bytes(T, L) <= 220 * 1024) return L;
  return 0;
}

bool band_dp_supported(int S, int T, int R) { return R <= kBandR && band_chunk_len(T) != 0; }

int launch_band_dp(const float *pxc, const float *pyc, const int32_t *ranges, const int32_t *boundary, int B, int S,
                   int T, int R, int rnnt_type, float delay_penalty, bool want_grad, void *workspace, float *gxc,
                   float *gyc, float *scores, cudaStream_t stream) {
  char *w = static_cast<char *>(workspace);
  const size_t nv = round_up_sz((size_t)B * (T + 1) * kBandR * sizeof(double), 256);
  const size_t no = round_up_sz((size_t)B * (T + 1) * sizeof(int), 256);
  BandDpParams p;
  p.pxc = pxc; p.pyc = pyc; p.ranges = ranges; p.boundary = boundary;
  p.va = reinterpret_cast<double *>(w); p.ub = reinterpret_cast<double *>(w + nv);
  p.oa = reinterpret_cast<int *>(w + 2 * nv); p.ob = reinterpret_cast<int *>(w + 2 * nv + no);
  p.img = reinterpret_cast<uint32_t *>(w + 2 * nv + 2 * no);
  p.img_frame = reinterpret_cast<int *>(w + 2 * nv + 2 * no + round_up_sz((size_t)B * 2 * (T + 1) * 64 * sizeof(uint32_t), 256));
  p.S = S; p.T = T; p.R = R; p.modified = (rnnt_type != FRN_REGULAR); p.rnnt_type = rnnt_type;
  p.delay_penalty = delay_penalty;
  p.L = band_chunk_len(T);
  if (R > kBandR || p.L == 0) return FRN_EUNSUPPORTED;
  const size_t smem = band_smem_bytes(T, p.L);
  cudaError_t e = cudaFuncSetAttribute(band_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return note_cuda_error(e);
  count_launch(), band_dp_kernel<<<dim3(B, want_grad ? 2 : 1), kBandThreads, smem, stream>>>(p);
  int rc = check_launch();
  if (rc) return rc;
  const int n = want_grad ? max(B * T * R, B) : B;
  count_launch(), band_finalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p, want_grad ? gxc : nullptr,
                                                                           want_grad ? gyc : nullptr, scores, B);
  return check_launch();
}

}  // namespace frn
