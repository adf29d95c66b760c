// A9: gradient of sum_b g_b * scores_b w.r.t. am and lm for rnnt_loss_simple —
// what TensorFlow autodiff produces through rnnt_loss.py:175-221 once
// _RNNTLossGrad (__init__.py:154-162) has supplied the occupation counts.
//
//   px[s,t] = am[t,sym_s] + lm[s,sym_s] - norm[s,t]      py[s,t] = am[t,blank] + lm[s,blank] - norm[s,t]
//   norm[s,t] = log Z[s,t] + lmmax[s] + ammax[t],  Z = sum_c lmp[s,c] amp[t,c] (+tiny)
//   G[s,t]  = gpy[s,t] + gpx[s,t]                         (total weight on -norm; constrained: gpx also feeds py[s+1,t])
//   W[s,t]  = G[s,t] / Z[s,t]
//   am_grad[t,c] = g ( [c=sym_s] sum_s gpx[s,t] + [c=blank] sum_s gpy[s,t] - amp[t,c] sum_s W[s,t] lmp[s,c] )
//   lm_grad[s,c] = g ( [c=sym_s] sum_t gpx[s,t] + [c=blank] sum_t gpy[s,t] - lmp[s,c] sum_t W[s,t] amp[t,c] )
//
// Smoothed loss (rnnt_loss.py:1266-1365), out = comb x + lm_scale x_lmonly + am_scale x_amonly:
//   the terms above get the factor comb (contractions) resp. comb + am_scale / comb + lm_scale (scatters), and
//   am_grad[t,c] -= g am_scale q[t,c] Gt[t],          q = amp u / D,  D[t] = sum_c amp[t,c] u[c],  Gt = sum_s G
//   lm_grad[s,c] -= g lm_scale r[s,c] Gs[s],          r = softmax(lm[s,:]),                        Gs = sum_t G
//   and, through the batch-global unigram u = mean_{b,s} r + tiny (rnnt_loss.py:1279-1280):
//   du[c] = sum_b g_b am_scale ( (sum_s Sx[s] [c=sym_s] + [c=blank] sum_s Sy[s]) / u[c] - sum_t Gt[t]/D[t] amp[t,c] )
//   lm_grad[b,s,c] += r[b,s,c] (du[c] - sum_c' r[b,s,c'] du[c']) / (B (S+1))
//   (Sx[s] = sum_t gpx[s,t], Sy[s] = sum_t gpy[s,t]; checked against finite differences of the float64 oracle,
//   tests/test_oracle_properties.py).
//
// First version: exact-FP32 SIMT tiles for the two contractions ([T x S1].[S1 x C]
// and [S1 x T].[T x C]); the tcgen05 version follows the forward kernel's scheme.
#include "common.cuh"
#include "launchers.h"
#include "simple_bwd_params.cuh"

namespace frn {

constexpr int kDuRows = 64;                 // am rows per block of the du partial sums

// Weight on px[b,s,t].  Regular lattice: the forward overwrote frame t_end of every px row with -inf
// (fix_for_boundary, rnnt_loss.py:51-60), a constant - whatever the caller put there passes nothing back
// (occupation counts are 0 there anyway; arbitrary cotangents of get_rnnt_logprobs are not).
__device__ __forceinline__ float gpx_at(const BwdParams &p, int b, int s, int t) {
  if (p.rnnt_type == FRN_REGULAR && t == p.boundary[4 * b + 3]) return 0.f;
  return p.gpx[((size_t)b * p.S + s) * p.T1 + t];
}

// W[b,s,t] = G / Z with Z = exp(norm - lmmax - ammax), norm = am[t,blank] + lm[s,blank] - py[s,t]; written over
// the padded [S1p][Tp] domain (zeros outside the lattice) so that the contraction tiles need no bounds checks
__global__ void __launch_bounds__(256) bwd_weights_kernel(BwdParams p) {
  const int S1 = p.S + 1;
  const size_t n = (size_t)p.B * p.S1p * p.Tp;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / ((size_t)p.S1p * p.Tp));
  const int rem = (int)(i - (size_t)b * p.S1p * p.Tp);
  const int s = rem / p.Tp, t = rem - s * p.Tp;
  float w = 0.f;
  if (s < S1 && t < p.T) {
    const size_t at = ((size_t)b * S1 + s) * p.T + t;
    float G = p.gpy[at];
    if (s < p.S) G += gpx_at(p, b, s, t);
    if (p.rnnt_type == FRN_CONSTRAINED && s >= 1) G += gpx_at(p, b, s - 1, t);  // px[s-1,t] += py[s,t]
    if (G != 0.f) {
      const float norm = p.am[((size_t)b * p.T + t) * p.C + p.term] + p.lm[((size_t)b * S1 + s) * p.C + p.term] - p.py[at];
      const float logZ = norm - p.lmmax[(size_t)b * S1 + s] - p.ammax[(size_t)b * p.T + t];
      w = G * expf(-logZ);
    }
  }
  p.W[i] = w;
}

// out[m,n] = -g * probs(x[m,n]) * sum_k Wk[m,k] * probs(y[k,n])   (+ scatter terms added afterwards)
// AM side: m = t, k = s, x = am, y = lm, Wk[m,k] = W[k][m];  LM side: m = s, k = t, x = lm, y = am, Wk[m,k] = W[m][k].
template <bool AM_SIDE>
__global__ void __launch_bounds__(256) bwd_contract_kernel(BwdParams p) {
  constexpr int TILE = 64, BK = 16;
  __shared__ float Ws[BK][TILE + 4];   // [k][m]
  __shared__ float Ys[BK][TILE + 4];   // [k][n] probabilities
  const int b = blockIdx.z, m0 = blockIdx.y * TILE, n0 = blockIdx.x * TILE;
  const int S1 = p.S + 1, C = p.C, T = p.T;
  const int M = AM_SIDE ? T : S1, K = AM_SIDE ? S1 : T;
  const float *x = (AM_SIDE ? p.am + (size_t)b * T * C : p.lm + (size_t)b * S1 * C);
  const float *y = (AM_SIDE ? p.lm + (size_t)b * S1 * C : p.am + (size_t)b * T * C);
  const float *xmax = AM_SIDE ? p.ammax + (size_t)b * T : p.lmmax + (size_t)b * S1;
  const float *ymax = AM_SIDE ? p.lmmax + (size_t)b * S1 : p.ammax + (size_t)b * T;
  const float *Wb = p.W + (size_t)b * p.S1p * p.Tp;
  const int Tp = p.Tp;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = tid; i < BK * TILE; i += 256) {
      const int kk = i / TILE, mm = i - kk * TILE;      // mm fastest
      const int k = k0 + kk, m = m0 + mm;
      float w = 0.f;
      if (k < K && m < M) w = AM_SIDE ? Wb[(size_t)k * Tp + m] : Wb[(size_t)m * Tp + k];
      Ws[kk][mm] = w;
      const int n = n0 + mm;
      float pr = 0.f;
      if (k < K && n < C) pr = expf(y[(size_t)k * C + n] - ymax[k]);
      Ys[kk][mm] = pr;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4 *>(&Ws[kk][ty * 4]);
      const float4 c = *reinterpret_cast<const float4 *>(&Ys[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float g = p.scores_grad ? p.scores_grad[b] : 1.f;
  float *out = AM_SIDE ? p.am_grad + (size_t)b * T * C : p.lm_grad + (size_t)b * S1 * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < C) out[(size_t)m * C + n] = -g * p.comb * expf(x[(size_t)m * C + n] - xmax[m]) * acc[i][j];
    }
  }
}

// Scatter terms of am_grad:  am_grad[t, sym_s] += g gpx[s,t],  am_grad[t, blank] += g sum_s gpy[s,t].
// Block = 32 consecutive frames of one utterance x 8 warps: lanes along t (the occupation counts are read
// coalesced), warp w takes the symbols s = w, w + 8, ...  A class can occur several times in an utterance: all its
// occurrences are summed, in a fixed order, by the thread holding the LAST one, so every address of a row
// receives exactly one addition from the symbol terms and one from the blank term - as fire-and-forget
// red.global.add, no read-modify-write latency chain (the first version walked the 101 symbols of a row with
// dependent load-add-store round trips: 89 us at the c2 shape).  Deterministic unless a symbol equals the blank id.
__global__ void __launch_bounds__(256) bwd_scatter_am_kernel(BwdParams p) {
  extern __shared__ int32_t sc_smem[];          // sym[S], prev[S] (previous occurrence or -1), tail[S]
  __shared__ float blank_part[8][32];
  const int S = p.S, S1 = p.S + 1;
  int32_t *s_sym = sc_smem, *s_prev = sc_smem + S, *s_tail = sc_smem + 2 * S;
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  const int32_t *sym = p.symbols + (size_t)b * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s_sym[i] = sym[i];
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const int c = s_sym[i];
    int prev = -1, tail = 1;
    for (int j = i - 1; j >= 0; --j)
      if (s_sym[j] == c) { prev = j; break; }
    for (int j = i + 1; j < S; ++j)
      if (s_sym[j] == c) { tail = 0; break; }
    s_prev[i] = prev; s_tail[i] = tail;
  }
  __syncthreads();
  const bool t_ok = t < p.T;
  const float g = (p.scores_grad ? p.scores_grad[b] : 1.f) * (p.smoothed ? p.comb + p.am_scale : 1.f);
  float *row = p.am_grad + ((size_t)b * p.T + (t_ok ? t : 0)) * p.C;
  float blank = 0.f;
  for (int s = w; s < S1; s += 8) {
    if (t_ok) blank += p.gpy[((size_t)b * S1 + s) * p.T + t];
    if (s < S && t_ok) {
      const float gx = gpx_at(p, b, s, t);
      if (p.rnnt_type == FRN_CONSTRAINED) blank += gx;   // px[s,t] also contains py[s+1,t]
      if (s_tail[s]) {
        float v = gx;
        for (int j = s_prev[s]; j >= 0; j = s_prev[j]) v += gpx_at(p, b, j, t);
        atomicAdd(row + s_sym[s], g * v);
      }
    }
  }
  blank_part[w][lane] = blank;
  __syncthreads();
  if (w == 0 && t_ok) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += blank_part[j][lane];
    atomicAdd(row + p.term, g * tot);
  }
}

__global__ void __launch_bounds__(256) bwd_scatter_lm_kernel(BwdParams p) {
  // one warp per (b,s): row sums over t
  const int S1 = p.S + 1;
  const int bs = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bs >= p.B * S1) return;
  const int b = bs / S1, s = bs - b * S1;
  float sx = 0.f, sy = 0.f;
  for (int t = lane; t < p.T; t += 32) {
    sy += p.gpy[(size_t)bs * p.T + t];
    if (s < p.S) sx += gpx_at(p, b, s, t);
    if (p.rnnt_type == FRN_CONSTRAINED && s >= 1) sy += gpx_at(p, b, s - 1, t);
  }
  sx = warp_sum(sx); sy = warp_sum(sy);
  const float g0 = p.scores_grad ? p.scores_grad[b] : 1.f;
  float *row = p.lm_grad + (size_t)bs * p.C;
  if (p.smoothed) {
    // lm-only term and the unigram path: both need the softmax of this lm row
    const float *x = p.lm + (size_t)bs * p.C;
    const float mx = p.lmmax[bs], inv = 1.f / p.lmsum[bs];
    float dot = 0.f;
    for (int c = lane; c < p.C; c += 32) dot += expf(x[c] - mx) * inv * p.du[c];
    dot = warp_sum(dot);
    const float invN = p.usums ? 1.f / p.usums[p.C] : 1.f / ((float)p.B * (float)S1);   // rows of the GLOBAL batch
    const float k = g0 * p.lm_scale * (sx + sy);
    for (int c = lane; c < p.C; c += 32) {
      const float r = expf(x[c] - mx) * inv;
      row[c] += r * ((p.du[c] - dot) * invN - k);
    }
    __syncwarp();
  }
  if (lane == 0) {
    const float g = g0 * (p.smoothed ? p.comb + p.lm_scale : 1.f);
    if (s < p.S) row[p.symbols[(size_t)b * p.S + s]] += g * sx;
    row[p.term] += g * sy;
  }
}

// ---- smoothed loss: column / row sums of the upstream weights ----
__global__ void __launch_bounds__(128) bwd_gt_kernel(BwdParams p) {
  const int S1 = p.S + 1;
  const int bt = blockIdx.x * blockDim.x + threadIdx.x;
  if (bt >= p.B * p.T) return;
  const int b = bt / p.T, t = bt - b * p.T;
  float acc = 0.f;
  for (int s = 0; s < S1; ++s) {
    acc += p.gpy[((size_t)b * S1 + s) * p.T + t];
    if (s < p.S) acc += gpx_at(p, b, s, t) * (p.rnnt_type == FRN_CONSTRAINED ? 2.f : 1.f);
  }
  p.Gt[bt] = acc;     // constrained: gpx[s,t] weighs px[s,t] and, through the fold, py[s+1,t]
}
__global__ void __launch_bounds__(256) bwd_rowsums_kernel(BwdParams p) {
  const int S1 = p.S + 1;
  const int bs = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bs >= p.B * S1) return;
  const int b = bs / S1, s = bs - b * S1;
  float sx = 0.f, sy = 0.f;
  for (int t = lane; t < p.T; t += 32) {
    sy += p.gpy[(size_t)bs * p.T + t];
    if (s < p.S) sx += gpx_at(p, b, s, t);
    if (p.rnnt_type == FRN_CONSTRAINED && s >= 1) sy += gpx_at(p, b, s - 1, t);
  }
  sx = warp_sum(sx); sy = warp_sum(sy);
  if (lane == 0) { p.Sx[bs] = sx; p.Sy[bs] = sy; }
}
// am-only term of am_grad, and the partial sums over am rows of  coef[b,t] amp[b,t,c]  for du
__global__ void __launch_bounds__(256) bwd_smooth_am_kernel(BwdParams p) {
  __shared__ float coef[kDuRows], mx[kDuRows];
  const int row0 = blockIdx.x * kDuRows, rows = min(kDuRows, p.B * p.T - row0);
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const int bt = row0 + i, b = bt / p.T;
    const float g = p.scores_grad ? p.scores_grad[b] : 1.f;
    mx[i] = p.ammax[bt];
    coef[i] = g * p.am_scale * p.Gt[bt] * expf(p.ammax[bt] - p.amonly[bt]);      // g am_scale Gt / D
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const float u = p.unigram[c];
    float acc = 0.f;
    for (int i = 0; i < rows; ++i) {
      const size_t at = (size_t)(row0 + i) * p.C + c;
      const float v = coef[i] * expf(p.am[at] - mx[i]);
      acc += v;
      p.am_grad[at] -= v * u;
    }
    p.partial[(size_t)blockIdx.x * p.C + c] = acc;
  }
}
// du[c]: one thread per class, sequential (deterministic) sums
__global__ void __launch_bounds__(128) bwd_du_kernel(BwdParams p, int chunks) {
  const int S1 = p.S + 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  float hist = 0.f;
  for (int b = 0; b < p.B; ++b) {
    const float g = (p.scores_grad ? p.scores_grad[b] : 1.f) * p.am_scale;
    float h = 0.f;
    for (int s = 0; s < p.S; ++s)
      if (p.symbols[(size_t)b * p.S + s] == c) h += p.Sx[(size_t)b * S1 + s];
    if (c == p.term)
      for (int s = 0; s < S1; ++s) h += p.Sy[(size_t)b * S1 + s];
    hist += g * h;
  }
  float v = 0.f;
  for (int k = 0; k < chunks; ++k) v += p.partial[(size_t)k * p.C + c];
  p.du[c] = hist / p.unigram[c] - v;
}

size_t simple_bwd_workspace_bytes(int B, int S, int T, int C) {
  const int T1 = T + 1;
  const size_t chunks = ((size_t)B * T + kDuRows - 1) / kDuRows;
  return round_up_sz((size_t)B * S * T1 * sizeof(float), 256) + round_up_sz((size_t)B * (S + 1) * T * sizeof(float), 256) +
         round_up_sz((size_t)B * round_up(S + 1, 128) * round_up(T, 128) * sizeof(float), 256) +
         simple_stats_bytes(B, S, T, C) + round_up_sz((size_t)B * T * sizeof(float), 256) +
         2 * round_up_sz((size_t)B * (S + 1) * sizeof(float), 256) + round_up_sz((size_t)C * sizeof(float), 256) +
         round_up_sz(chunks * C * sizeof(float), 256);
}

// `unigram_sums` / `du_io` / `phase`: batch sharded by utterance.  The unigram is a function of
// every rank's lm, so d loss / d unigram (du, [C]) has to be summed over the ranks before it flows back into the
// lm rows: phase 1 runs everything up to this rank's share of du (written to du_io), the caller all-reduces
// du_io, phase 2 (same workspace, same arguments) adds the unigram and lm-only terms to lm_grad.  phase 0: one
// rank, everything.
int launch_simple_bwd(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                      const float *px_grad, const float *py_grad, const float *scores_grad, int B, int S, int T, int C,
                      int term, int rnnt_type, int smoothed, float lm_only_scale, float am_only_scale, float *am_grad,
                      float *lm_grad, void *workspace, cudaStream_t stream, const float *unigram_sums, float *du_io,
                      int phase) {
  const int S1 = S + 1, T1 = (rnnt_type == FRN_REGULAR) ? T + 1 : T;
  const int chunks = (B * T + kDuRows - 1) / kDuRows;
  char *w = static_cast<char *>(workspace);
  float *px = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S * (T + 1) * sizeof(float), 256);
  float *py = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * T * sizeof(float), 256);
  const int S1p = round_up(S1, 128), Tp = round_up(T, 128);
  float *W = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1p * Tp * sizeof(float), 256);
  void *stats = w; w += simple_stats_bytes(B, S, T, C);
  float *Gt = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
  float *Sx = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *Sy = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *du = reinterpret_cast<float *>(w); w += round_up_sz((size_t)C * sizeof(float), 256);
  float *partial = reinterpret_cast<float *>(w);
  if (du_io) du = du_io;
  int rc = FRN_OK;
  if (phase != 2) {
  // forward log-probs again (py gives Z); non-smoothed, and without the constrained px += py fold
  rc = launch_simple_logprobs(lm, am, symbols, boundary, B, S, T, C, term,
                              rnnt_type == FRN_CONSTRAINED ? FRN_MODIFIED : rnnt_type, 0, 0.f, 0.f, px, py, stats,
                              stream);
  if (rc) return rc;
  if (smoothed) {   // row sums, unigram and am-only normalisers into the same statistics block
    rc = launch_smoothing_stats(lm, am, B, S, T, C, stats, stream, unigram_sums);
    if (rc) return rc;
  }
  }
  char *sw = static_cast<char *>(stats);
  BwdParams p;
  p.lm = lm; p.am = am; p.symbols = symbols; p.boundary = boundary; p.gpx = px_grad; p.gpy = py_grad; p.py = py;
  p.lmmax = reinterpret_cast<float *>(sw);
  p.lmsum = reinterpret_cast<float *>(sw + round_up_sz((size_t)B * S1 * sizeof(float), 256));
  p.ammax = reinterpret_cast<float *>(sw + 2 * round_up_sz((size_t)B * S1 * sizeof(float), 256));
  p.amonly = reinterpret_cast<float *>(sw + 2 * round_up_sz((size_t)B * S1 * sizeof(float), 256) +
                                       round_up_sz((size_t)B * T * sizeof(float), 256));
  p.unigram = reinterpret_cast<float *>(sw + 2 * round_up_sz((size_t)B * S1 * sizeof(float), 256) +
                                        2 * round_up_sz((size_t)B * T * sizeof(float), 256));
  p.scores_grad = scores_grad; p.W = W; p.S1p = S1p; p.Tp = Tp; p.am_grad = am_grad; p.lm_grad = lm_grad;
  p.B = B; p.S = S; p.T = T; p.T1 = T1; p.C = C; p.term = term; p.rnnt_type = rnnt_type;
  p.smoothed = smoothed;
  const double lms = (double)lm_only_scale, ams = (double)am_only_scale;      // rnnt_loss.py:1342-1349
  p.comb = smoothed ? (float)(1.0 - lms - ams) : 1.f;
  p.lm_scale = (float)(lms == 0.0 ? 1.0e-20 : lms);
  p.am_scale = (float)(ams == 0.0 ? 1.0e-20 : ams);
  p.Gt = Gt; p.Sx = Sx; p.Sy = Sy; p.du = du; p.partial = partial; p.usums = unigram_sums;
  if (phase == 2) {
    count_launch(), bwd_scatter_lm_kernel<<<(B * S1 + 7) / 8, 256, 0, stream>>>(p);
    return check_launch();
  }
  const size_t n = (size_t)B * S1p * Tp;
  count_launch(), bwd_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p);
  // the two contractions: tcgen05 (simple_bwd_tc.cu); the exact-FP32 SIMT tiles serve shapes it cannot take
  // (C % 4 != 0) and FRN_BWD_SIMT=1 forces them in the debug-hooks build (cross-check)
  rc = debug_env_int("FRN_BWD_SIMT", 0) == 1 ? FRN_EUNSUPPORTED : launch_bwd_contract_tc(p, stream);
  if (rc == FRN_EUNSUPPORTED) {
    dim3 g_am((C + 63) / 64, (T + 63) / 64, B), g_lm((C + 63) / 64, (S1 + 63) / 64, B);
    count_launch(), bwd_contract_kernel<true><<<g_am, 256, 0, stream>>>(p);
    count_launch(), bwd_contract_kernel<false><<<g_lm, 256, 0, stream>>>(p);
  } else if (rc) {
    return rc;
  }
  if (smoothed) {
    count_launch(), bwd_gt_kernel<<<(B * T + 127) / 128, 128, 0, stream>>>(p);
    count_launch(), bwd_rowsums_kernel<<<(B * S1 + 7) / 8, 256, 0, stream>>>(p);
    count_launch(), bwd_smooth_am_kernel<<<chunks, 256, 0, stream>>>(p);
    count_launch(), bwd_du_kernel<<<(C + 127) / 128, 128, 0, stream>>>(p, chunks);
  }
  const size_t sc_bytes = 3 * (size_t)S * sizeof(int32_t);
  if (sc_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(bwd_scatter_am_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_bytes);
    if (e != cudaSuccess) return note_cuda_error(e);
  }
  count_launch(), bwd_scatter_am_kernel<<<dim3((T + 31) / 32, B), 256, sc_bytes, stream>>>(p);
  if (phase != 1) count_launch(), bwd_scatter_lm_kernel<<<(B * S1 + 7) / 8, 256, 0, stream>>>(p);
  return check_launch();
}

}  // namespace frn
