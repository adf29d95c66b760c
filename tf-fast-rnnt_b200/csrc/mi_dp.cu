// Lattice recursion ("mutual information recursion") for sm_100a.
//
// Replaces the reference's K1/K2 (tf_fast_rnnt/csrc/mutual_information_cuda.cu:
// 174-422 forward, 490-760 backward, launched 2*(S/32+T/32+1) times from host
// loops at :799-808 and :860-872) with three launches:
//
//   skew     px/py [B][S][T1] / [B][S+1][T]  ->  X/Y [B][d][s'] (diagonal-major,
//            log2 domain, boundary masks applied, -inf replaced by a finite
//            sentinel).  Fully parallel, HBM/L2 bound.
//   chain    one CTA per (utterance, direction), a software pipeline of warps
//            (see dp_chain_kernel); one step = one anti-diagonal (d = t' + k s',
//            k = 1 regular / 0 modified), neighbours exchanged by warp shuffle, the
//            diagonal-major arc scores streamed into a shared-memory ring by
//            1-D bulk async copies (TMA engine) behind mbarriers.  Forward
//            (alpha) and backward (beta) chains run concurrently in different
//            CTAs.  Each row keeps an exact integer offset plus a small float32
//            residual, re-centred every step off the dependency chain (see the
//            comment above ChainParams), which makes the float32 result two
//            orders of magnitude closer to float64 truth than the reference's
//            plain float32 p[].
//            Latency bound: (S_b + T_b) dependent log-adds.
//   finalize occupation counts  px_grad = exp(alpha + px + beta' - total),
//            py_grad likewise (equal to the reference's p_grad recursion,
//            cu:472-481, in exact arithmetic), written in the reference layout.
//
#include "common.cuh"

namespace frn {

constexpr int kStages = 3;

// ---------------------------------------------------------------------------
// skew (dense input)
// ---------------------------------------------------------------------------
struct SkewDenseParams {
  const float *px, *py;     // reference layout
  const int32_t *boundary;  // [B][4]
  float *X, *Y;             // [B][Dn][P]
  int S, T, T1, P, Dn;
  float delay_penalty;      // added to px (rnnt_loss.py:316-321); 0 = none
};

// One block: 32 diagonals x 32 rows of one utterance.
template <int K>
__global__ void __launch_bounds__(256) skew_dense_kernel(SkewDenseParams p) {
  constexpr int kPitch = 65 - K;
  constexpr int kWidth = 32 + 31 * K;  // t' values touched by the tile
  __shared__ float sx[32 * kPitch], sy[32 * kPitch];
  const int b = blockIdx.z;
  const int d0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const int off = K ? 0 : -1;
  const int tlo = d0 - K * (s0 + 31);  // smallest t' of the tile
  const float *pxb = p.px + (size_t)b * p.S * p.T1;
  const float *pyb = p.py + (size_t)b * (p.S + 1) * p.T;

  // load: row = s' - s0, col = t' - tlo, col fastest (coalesced along t)
  for (int i = threadIdx.x; i < 32 * kWidth; i += blockDim.x) {
    const int row = i / kWidth, col = i - row * kWidth;
    const int sp = s0 + row, tp = tlo + col;  // destination cell (s', t')
    float vx = kNeg, vy = kNeg;
    if (sp <= Sb && tp <= Tb) {
      // mutual_information_cuda.cu:295-303 (forward load rules)
      if (sp >= 1 && tp + off >= 0) {
        const int t_abs = t_begin + tp + off;
        float v = pxb[(size_t)(s_begin + sp - 1) * p.T1 + t_abs];
        if (p.delay_penalty != 0.f) v += delay_penalty_value(bd.w, t_abs, p.delay_penalty);
        vx = fmaxf(v * kLog2e, kNeg);
      }
      if (tp >= 1) vy = fmaxf(pyb[(size_t)(s_begin + sp) * p.T + t_begin + tp - 1] * kLog2e, kNeg);
    }
    sx[row * kPitch + col] = vx;
    sy[row * kPitch + col] = vy;
  }
  __syncthreads();
  // store: s' fastest (coalesced along the row axis of X/Y)
  float *Xb = p.X + (size_t)b * p.Dn * p.P, *Yb = p.Y + (size_t)b * p.Dn * p.P;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int dd = i >> 5, ss = i & 31;
    const int d = d0 + dd;
    if (d >= p.Dn) continue;
    const int col = dd + K * (31 - ss);
    Xb[(size_t)d * p.P + s0 + ss] = sx[ss * kPitch + col];
    Yb[(size_t)d * p.P + s0 + ss] = sy[ss * kPitch + col];
  }
}

// ---------------------------------------------------------------------------
// chain
// ---------------------------------------------------------------------------
// Numerics.  Every lattice value is carried as  o + r : `o` an exact integer
// (held in a float, |o| < 2^24) and `r` a small float32 residual.  At every
// step each row first moves k = rint(r_previous) from its residual into its
// offset (both operations are exact in float32; k depends on the previous
// step only, so none of this is on the dependency chain), a row that is still
// "minus infinity" adopts the offset of the row that feeds it, and the exact
// offset difference between neighbouring rows is folded into the arc score.
// All float32 roundings therefore happen at magnitude ~10 instead of |p| ~ 10^3
// as in the reference's plain float32 p[]: measured 4e-6 max relative error on
// the occupation counts at the c2 shape against 1.3e-3 for the reference's
// arithmetic ("numerics").
struct ChainParams {
  const float *X, *Y;
  float *ar, *ao;      // [B][Dn][P]  forward residual / offset (dir 0)
  float *bx, *by, *bo; // [B][Dn][P]  backward-side operands and their frame offset (dir 1)
  const int32_t *boundary;
  int k, P, Dn, S, T;
  int CH, NST;         // diagonals per bulk copy (divides kChunk); ring stages (>= warps + 2)
};

// Execution.  One CTA per (utterance, direction); RPL consecutive lattice rows
// per lane, 32*RPL rows per warp, W = P / (32*RPL) warps.  The warps form a
// software pipeline: warp w runs one chunk (CH diagonals) behind the warp that
// owns the rows feeding it and picks the boundary row's (residual, offset) of
// every step out of a shared-memory ring that warp filled; one mbarrier per
// (warp, ring slot) says "chunk published".  There is no per-step block barrier
// and every warp sits alone on its SM sub-partition, so a step costs one
// dependent log-add (shuffle + ex2 + lg2 + a few adds) rather than the issue
// time of four interleaved rows.  X/Y chunks are shared by all warps: NST >= W+2
// stages of 1-D bulk copies; the tail warp of the pipeline recycles a stage.
template <int RPL, int DIR>
__device__ __forceinline__ void dp_chain_body(const ChainParams &p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5;
  const int P = p.P, CH = p.CH, NST = p.NST;
  const int stage_floats = 2 * CH * P;
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *mbar_xy = reinterpret_cast<uint64_t *>(ring + NST * stage_floats);
  uint64_t *mbar_edge = mbar_xy + NST;                               // [W][NST], indexed by consumer warp
  float2 *edge = reinterpret_cast<float2 *>(mbar_edge + W * NST);    // [W][NST][CH], indexed by consumer warp

  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  if (Sb < 0 || Tb < 0 || bd.x < 0 || bd.y < 0 || bd.z > p.S || bd.w > p.T) return;  // finalize reports it
  const int Db = Tb + p.k * Sb;
  const int nchunk = Db / CH + 1;
  const size_t plane = (size_t)b * p.Dn * P;
  const float *Xg = p.X + plane, *Yg = p.Y + plane;
  const uint32_t chunk_bytes = (uint32_t)(CH * P * sizeof(float));

  auto issue = [&](int seq, int st) {
    const int ci = DIR ? nchunk - 1 - seq : seq;
    mbar_arrive_expect_tx(&mbar_xy[st], 2 * chunk_bytes);
    bulk_g2s(ring + st * stage_floats, Xg + (size_t)ci * CH * P, chunk_bytes, &mbar_xy[st]);
    bulk_g2s(ring + st * stage_floats + CH * P, Yg + (size_t)ci * CH * P, chunk_bytes, &mbar_xy[st]);
  };

  if (tid == 0) {
    for (int i = 0; i < NST + W * NST; ++i) mbar_init(&mbar_xy[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < NST && i < nchunk; ++i) issue(i, i);

  const int pos = DIR ? (W - 1 - w) : w;        // position in the warp pipeline, 0 = head
  const bool is_tail = (pos == W - 1), fed = (pos > 0), feeds = (pos < W - 1);
  const int wc = DIR ? w - 1 : w + 1;           // the warp this one feeds
  const int r0 = RPL * tid;                     // first lattice row of this lane
  const bool lane_in = DIR ? (lane == 31) : (lane == 0);    // lane fed across the warp boundary
  const bool lane_out = DIR ? (lane == 0) : (lane == 31);   // lane feeding the next warp
  const bool global_edge = DIR ? (tid == (int)blockDim.x - 1) : (tid == 0);  // row with no neighbour at all
  const bool take_edge = lane_in && fed, publish = lane_out && feeds;
  constexpr int step_sign = DIR ? -1 : 1;

  float r[RPL], o[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    r[j] = ((r0 + j) == (DIR ? Sb : 0)) ? 0.f : kNeg;
    o[j] = 0.f;
  }
  float *const out0 = (DIR ? p.bx : p.ar) + plane + r0;   // alpha: residual | beta: px-arc operand
  float *const out1 = (DIR ? p.bo : p.ao) + plane + r0;   // frame offset
  float *const out2 = p.by + plane + r0;                  // beta only: py-arc operand
  if (!DIR) {
#pragma unroll
    for (int j = 0; j < RPL; ++j) { out0[j] = r[j]; out1[j] = 0.f; }
  }
  // state of the row feeding lane_in before the first step (initial condition)
  float2 carry = make_float2(kNeg, 0.f);
  if (DIR && (r0 + RPL) == Sb) carry.x = 0.f;

  for (int i = 0; i < nchunk; ++i) {
    const int st = i % NST;
    const uint32_t par = (uint32_t)((i / NST) & 1);
    const int ci = DIR ? nchunk - 1 - i : i;
    mbar_wait(&mbar_xy[st], par);
    if (fed) mbar_wait(&mbar_edge[w * NST + st], par);   // the feeding warp has published this chunk
    const float *xs = ring + st * stage_floats + r0, *ys = xs + CH * P;
    const float2 *ein = edge + (size_t)(w * NST + st) * CH;
    float2 *eout = edge + (size_t)((feeds ? wc : w) * NST + st) * CH;
    const int e_lo = max(ci * CH, 1), e_hi = min(ci * CH + CH - 1, Db);
    const int n = e_hi - e_lo + 1;
    int e = DIR ? e_hi : e_lo;
    int el = e - ci * CH;
    // operands of the first step of the chunk
    float x[RPL], y[RPL], xnext = kNeg;
    float2 ev = carry;
#pragma unroll
    for (int j = 0; j < RPL; ++j) { x[j] = kNeg; y[j] = kNeg; }
    if (n > 0) {
#pragma unroll
      for (int j = 0; j < RPL; ++j) { x[j] = xs[el * P + j]; y[j] = ys[el * P + j]; }
      if (DIR && r0 + RPL < P) xnext = xs[el * P + RPL];
    }
    for (int q = 0; q < n; ++q) {
      // ---- prefetch the operands of the next step (off the dependency chain) ----
      const int eln = (q + 1 < n) ? el + step_sign : el;
      float x2[RPL], y2[RPL], xnext2 = kNeg;
#pragma unroll
      for (int j = 0; j < RPL; ++j) { x2[j] = xs[eln * P + j]; y2[j] = ys[eln * P + j]; }
      if (DIR && r0 + RPL < P) xnext2 = xs[eln * P + RPL];
      const float2 ev2 = ein[el];   // the feeding row's state after ITS step e = input of our next step (broadcast read)

      // ---- neighbour across the lane boundary: raw residual + the frame it is expressed in ----
      float nb_r = DIR ? __shfl_down_sync(0xffffffffu, r[0], 1) : __shfl_up_sync(0xffffffffu, r[RPL - 1], 1);
      float nb_o = DIR ? __shfl_down_sync(0xffffffffu, o[0], 1) : __shfl_up_sync(0xffffffffu, o[RPL - 1], 1);
      nb_r = lane_in ? (fed ? ev.x : kNeg) : nb_r;
      nb_o = take_edge ? ev.y : nb_o;

      // ---- lag-1 re-centring (exact): k = rint(previous residual) moves into the offset ----
      bool alive[RPL];
      float kk[RPL], on[RPL], rc[RPL];
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        alive[j] = r[j] > kNegThresh;
        kk[j] = alive[j] ? rintf(r[j]) : 0.f;
        rc[j] = r[j] - kk[j];
      }
      // A dead row adopts the NEW frame of the row that feeds it (across the lane
      // boundary that frame is re-derived from the shuffled raw residual and old
      // offset), so a moving wave-front never inherits a stale frame.
      const float nb_k = (nb_r > kNegThresh) ? rintf(nb_r) : 0.f;
      if (!DIR) {
        // alpha_e(s') = logadd(alpha_{e-1}(s'-1) + X[e][s'], alpha_{e-1}(s') + Y[e][s'])
        on[0] = (alive[0] || global_edge) ? o[0] + kk[0] : nb_o + nb_k;
#pragma unroll
        for (int j = 1; j < RPL; ++j) on[j] = alive[j] ? o[j] + kk[j] : on[j - 1];
        float nn[RPL];
        nn[0] = logadd2((nb_r + x[0]) + (nb_o - on[0]), rc[0] + y[0]);
#pragma unroll
        for (int j = 1; j < RPL; ++j) nn[j] = logadd2(rc[j - 1] + (x[j] + (on[j - 1] - on[j])), rc[j] + y[j]);
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          r[j] = nn[j]; o[j] = on[j];
          out0[(unsigned)(e * P + j)] = r[j];
          out1[(unsigned)(e * P + j)] = o[j];
        }
      } else {
        // beta_{e-1}(s') = logadd(X[e][s'+1] + beta_e(s'+1), Y[e][s'] + beta_e(s'))
        on[RPL - 1] = (alive[RPL - 1] || global_edge) ? o[RPL - 1] + kk[RPL - 1] : nb_o + nb_k;
#pragma unroll
        for (int j = RPL - 2; j >= 0; --j) on[j] = alive[j] ? o[j] + kk[j] : on[j + 1];
        float a[RPL], c[RPL];
        a[RPL - 1] = (nb_r + xnext) + (nb_o - on[RPL - 1]);
#pragma unroll
        for (int j = 0; j < RPL - 1; ++j) a[j] = rc[j + 1] + (x[j + 1] + (on[j + 1] - on[j]));
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          c[j] = rc[j] + y[j];
          // operands of diagonal e-1, expressed in the frame `on`
          out0[(unsigned)((e - 1) * P + j)] = a[j];
          out2[(unsigned)((e - 1) * P + j)] = c[j];
          out1[(unsigned)((e - 1) * P + j)] = on[j];
          r[j] = logadd2(a[j], c[j]);
          o[j] = on[j];
        }
      }
      if (publish) eout[el] = DIR ? make_float2(r[0], o[0]) : make_float2(r[RPL - 1], o[RPL - 1]);
      // rotate the prefetched operands in
#pragma unroll
      for (int j = 0; j < RPL; ++j) { x[j] = x2[j]; y[j] = y2[j]; }
      xnext = xnext2;
      ev = ev2;
      e += step_sign;
      el = eln;
    }
    carry = ev;   // the feeding row's state after the last step of this chunk
    if (publish) mbar_arrive(&mbar_edge[wc * NST + st]);
    if (is_tail) {
      __syncwarp();
      if (lane == 0 && i + NST < nchunk) issue(i + NST, st);
    }
  }
}

// grid = (B, 2): blockIdx.y selects the direction so that the forward and the
// backward chain of every utterance run concurrently on different SMs.
template <int RPL>
__global__ void __launch_bounds__(256, 1) dp_chain_kernel(ChainParams p) {
  if (blockIdx.y == 0) dp_chain_body<RPL, 0>(p);
  else dp_chain_body<RPL, 1>(p);
}

// ---------------------------------------------------------------------------
// finalize (dense output)
// ---------------------------------------------------------------------------
struct FinalizeDenseParams {
  const float *ar, *ao, *bx, *by, *bo;
  const int32_t *boundary;
  float *ans;               // [B]
  float *px_grad, *py_grad; // reference layout, may be null
  int S, T, T1, P, Dn, k;
};

__device__ __forceinline__ bool boundary_ok(const int4 &bd, int S, int T) {
  return bd.z - bd.x >= 0 && bd.w - bd.y >= 0 && bd.x >= 0 && bd.y >= 0 && bd.z <= S && bd.w <= T;
}

__global__ void dp_ans_kernel(FinalizeDenseParams p, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  if (!boundary_ok(bd, p.S, p.T)) { p.ans[b] = 0.f; return; }
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const size_t at = ((size_t)b * p.Dn + Tb + p.k * Sb) * p.P + Sb;
  const float tr = p.ar[at], to = p.ao[at];
  p.ans[b] = (tr < kNegThresh) ? -INFINITY : (float)(((double)tr + (double)to) * 0.6931471805599453);
}

// One block: 32 absolute diagonals (dabs = t + K s) x 32 rows (s) of one
// utterance.  The planes are read exactly once, coalesced along s; results go
// through shared memory and leave coalesced along t in the reference layout.
// Every element of px_grad / py_grad is written (zeros outside the boundary
// box): this replaces the reference's two cudaMemsetAsync (op.cc:94-98).
template <int K>
__global__ void __launch_bounds__(256) finalize_dense_kernel(FinalizeDenseParams p) {
  __shared__ float sgx[32][33], sgy[32][33];
  const int b = blockIdx.z;
  const int d0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y;
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const bool ok = boundary_ok(bd, p.S, p.T);
  const int noff = K ? 0 : 1;
  const int Db = Tb + K * Sb;
  const size_t plane = (size_t)b * p.Dn * p.P;
  float tot_r = 0.f, tot_o = 0.f;
  bool dead = !ok;
  if (ok) {
    tot_r = p.ar[plane + (size_t)Db * p.P + Sb];
    tot_o = p.ao[plane + (size_t)Db * p.P + Sb];
    dead = tot_r < kNegThresh;
  }
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  // phase 1: lane <-> s (coalesced plane reads), 8 warps stride the diagonals
  for (int dd = wrp; dd < 32; dd += 8) {
    const int s = s0 + lane, t = d0 + dd - K * s;  // absolute cell
    float vx = 0.f, vy = 0.f;
    if (!dead && t >= 0) {
      const int sp = s - s_begin, tp = t - t_begin;
      if (sp >= 0 && tp >= 0 && sp <= Sb && tp <= Tb) {
        const int d = tp + K * sp;
        if (d < Db) {  // arcs leave diagonals 0..Db-1
          const size_t at = plane + (size_t)d * p.P + sp;
          const float base = (p.ar[at] - tot_r) + ((p.ao[at] + p.bo[at]) - tot_o);
          // arc (s,t)->(s+1,t+noff): cu:727-746; arc (s,t)->(s,t+1): cu:747-753
          if (sp < Sb && tp + noff <= Tb) vx = ex2_approx(p.bx[at] + base);
          if (tp < Tb) vy = ex2_approx(p.by[at] + base);
        }
      }
    }
    sgx[lane][dd] = vx;
    sgy[lane][dd] = vy;
  }
  __syncthreads();
  // phase 2: lane <-> t within a row segment
  float *gx = p.px_grad + (size_t)b * p.S * p.T1;
  float *gy = p.py_grad + (size_t)b * (p.S + 1) * p.T;
  for (int ss = wrp; ss < 32; ss += 8) {
    const int s = s0 + ss;
    if (s > p.S) continue;
    const int t = d0 + lane - K * s;
    if (t < 0) continue;
    if (s < p.S && t < p.T1) gx[(size_t)s * p.T1 + t] = sgx[ss][lane];
    if (t < p.T) gy[(size_t)s * p.T + t] = sgy[ss][lane];
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    p.ans[b] = !ok ? 0.f : (dead ? -INFINITY : (float)(((double)tot_r + (double)tot_o) * 0.6931471805599453));
}

// ---------------------------------------------------------------------------
// host-side launchers (used by api.cu)
// ---------------------------------------------------------------------------
int launch_skew_dense(const float *px, const float *py, const int32_t *boundary, const DpGeom &g,
                      const DpWorkspace &w, float delay_penalty, cudaStream_t stream) {
  SkewDenseParams sp{px, py, boundary, w.X, w.Y, g.S, g.T, g.T1, g.P, g.Dn, delay_penalty};
  dim3 grid((g.Dn + 31) / 32, g.P / 32, g.B);
  if (g.k) skew_dense_kernel<1><<<grid, 256, 0, stream>>>(sp);
  else skew_dense_kernel<0><<<grid, 256, 0, stream>>>(sp);
  return check_launch();
}

struct ChainConfig { int W, CH, NST; size_t smem; };
static ChainConfig chain_config(const DpGeom &g) {
  ChainConfig c;
  c.W = g.P / (32 * g.rpl);
  c.NST = c.W + 2;
  c.CH = kChunk;
  auto bytes = [&](int ch) {
    return (size_t)c.NST * 2 * ch * g.P * sizeof(float) + (size_t)(c.NST + c.W * c.NST) * sizeof(uint64_t) +
           (size_t)c.W * c.NST * ch * sizeof(float2) + 128;
  };
  while (c.CH > 1 && bytes(c.CH) > 200 * 1024) c.CH >>= 1;
  c.smem = bytes(c.CH);
  return c;
}

int launch_chain(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, bool both_directions,
                 cudaStream_t stream) {
  if (g.P > kMaxRowsDp) return FRN_EUNSUPPORTED;
  const ChainConfig c = chain_config(g);
  ChainParams cp{w.X, w.Y, w.ar, w.ao, w.bx, w.by, w.bo, boundary, g.k, g.P, g.Dn, g.S, g.T, c.CH, c.NST};
  dim3 grid(g.B, both_directions ? 2 : 1);
  const int threads = 32 * c.W;
  cudaError_t e;
#define FRN_LAUNCH_CHAIN(RPL_)                                                                              \
  e = cudaFuncSetAttribute(dp_chain_kernel<RPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem); \
  if (e != cudaSuccess) return note_cuda_error(e);                                                          \
  dp_chain_kernel<RPL_><<<grid, threads, c.smem, stream>>>(cp);
  if (g.rpl == 1) { FRN_LAUNCH_CHAIN(1) }
  else if (g.rpl == 2) { FRN_LAUNCH_CHAIN(2) }
  else { FRN_LAUNCH_CHAIN(4) }
#undef FRN_LAUNCH_CHAIN
  return check_launch();
}

int launch_finalize_dense(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, float *ans,
                          float *px_grad, float *py_grad, cudaStream_t stream) {
  FinalizeDenseParams fp{w.ar, w.ao, w.bx, w.by, w.bo, boundary, ans, px_grad, py_grad,
                         g.S, g.T, g.T1, g.P, g.Dn, g.k};
  if (px_grad == nullptr || py_grad == nullptr) {
    dp_ans_kernel<<<(g.B + 127) / 128, 128, 0, stream>>>(fp, g.B);
    return check_launch();
  }
  // absolute diagonals dabs = t + k*s for t in [0, T], s in [0, S]
  dim3 grid((g.T + 1 + g.k * g.S + 31) / 32, (g.S + 1 + 31) / 32, g.B);
  if (g.k) finalize_dense_kernel<1><<<grid, 256, 0, stream>>>(fp);
  else finalize_dense_kernel<0><<<grid, 256, 0, stream>>>(fp);
  return check_launch();
}

}  // namespace frn
