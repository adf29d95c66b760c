// Lattice recursion ("mutual information recursion") for sm_100a.
//
// Replaces the reference's K1/K2 (tf_fast_rnnt/csrc/mutual_information_cuda.cu:
// 174-422 forward, 490-760 backward, launched 2*(S/32+T/32+1) times from host
// loops at :799-808 and :860-872) with three launches:
//
//   skew     px/py [B][S][T1] / [B][S+1][T]  ->  X/Y [B][d][s'] (diagonal-major,
//            log2 domain, boundary masks applied, -inf replaced by a finite
//            sentinel).  Fully parallel, HBM/L2 bound.
//   chain    one CTA per (utterance, direction).  Each lane owns 4 consecutive
//            lattice rows; one step = one anti-diagonal (d = t' + k s', k = 1
//            regular / 0 modified), neighbours exchanged by warp shuffle, the
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
  int k, P, Dn, S, T, CH;  // CH: diagonals per bulk copy (divides kChunk)
};

template <bool MULTI>
__global__ void __launch_bounds__(256, 1) dp_chain_kernel(ChainParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x, dir = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5;
  const int P = p.P, CH = p.CH;
  const int stage_floats = 2 * CH * P;
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(ring + kStages * stage_floats);
  float *edge_r = reinterpret_cast<float *>(mbar + kStages);  // [2][8] residual hand-off between warps
  float *edge_o = edge_r + 16;                                // [2][8] offset hand-off

  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  if (Sb < 0 || Tb < 0 || bd.x < 0 || bd.y < 0 || bd.z > p.S || bd.w > p.T) return;  // finalize reports it
  const int Db = Tb + p.k * Sb;
  const int nchunk = Db / CH + 1;
  const size_t plane = (size_t)b * p.Dn * P;
  const float *Xg = p.X + plane, *Yg = p.Y + plane;
  float *outAr = p.ar + plane, *outAo = p.ao + plane;
  float *outBx = p.bx + plane, *outBy = p.by + plane, *outBo = p.bo + plane;
  const uint32_t chunk_bytes = (uint32_t)(CH * P * sizeof(float));

  auto issue = [&](int seq, int st) {
    const int ci = dir ? nchunk - 1 - seq : seq;
    mbar_arrive_expect_tx(&mbar[st], 2 * chunk_bytes);
    bulk_g2s(ring + st * stage_floats, Xg + (size_t)ci * CH * P, chunk_bytes, &mbar[st]);
    bulk_g2s(ring + st * stage_floats + CH * P, Yg + (size_t)ci * CH * P, chunk_bytes, &mbar[st]);
  };

  if (tid == 0) {
    for (int st = 0; st < kStages; ++st) mbar_init(&mbar[st], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < kStages && i < nchunk; ++i) issue(i, i);

  const int r0 = kRowsPerLane * tid;  // first lattice row of this lane
  const bool first_row_lane = (tid == 0), last_row_lane = (tid == (int)blockDim.x - 1);
  float r[4], o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r[j] = ((r0 + j) == (dir ? Sb : 0)) ? 0.f : kNeg;
    o[j] = 0.f;
  }
  if (!dir) {
    *reinterpret_cast<float4 *>(outAr + r0) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4 *>(outAo + r0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (MULTI) {
    if (lane == (dir ? 0 : 31)) {
      edge_r[8 + w] = dir ? r[0] : r[3];
      edge_o[8 + w] = 0.f;
    }
    __syncthreads();
  }

  int step = 0;
  for (int i = 0; i < nchunk; ++i) {
    const int st = i % kStages;
    const int ci = dir ? nchunk - 1 - i : i;
    mbar_wait(&mbar[st], (uint32_t)((i / kStages) & 1));
    const float *xs = ring + st * stage_floats, *ys = xs + CH * P;
    const int e_lo = max(ci * CH, 1), e_hi = min(ci * CH + CH - 1, Db);
    const int n = e_hi - e_lo + 1;
    float4 x = make_float4(kNeg, kNeg, kNeg, kNeg), y = x;
    float xn = kNeg;
    if (n > 0) {
      const int el = (dir ? e_hi : e_lo) - ci * CH;
      x = *reinterpret_cast<const float4 *>(xs + el * P + r0);
      y = *reinterpret_cast<const float4 *>(ys + el * P + r0);
      if (dir && r0 + 4 < P) xn = xs[el * P + r0 + 4];
    }
    for (int q = 0; q < n; ++q) {
      const int e = dir ? e_hi - q : e_lo + q;
      // prefetch the next diagonal's arc scores (off the dependency chain)
      float4 x2 = x, y2 = y;
      float xn2 = kNeg;
      if (q + 1 < n) {
        const int el2 = (dir ? e - 1 : e + 1) - ci * CH;
        x2 = *reinterpret_cast<const float4 *>(xs + el2 * P + r0);
        y2 = *reinterpret_cast<const float4 *>(ys + el2 * P + r0);
        if (dir && r0 + 4 < P) xn2 = xs[el2 * P + r0 + 4];
      }
      // lag-1 re-centring (exact): k = rint(previous residual) moves into the offset
      bool alive[4];
      float kk[4], on[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        alive[j] = r[j] > kNegThresh;
        kk[j] = alive[j] ? rintf(r[j]) : 0.f;
      }
      if (!dir) {
        // alpha_e(s') = logadd(alpha_{e-1}(s'-1) + X[e][s'], alpha_{e-1}(s') + Y[e][s'])
        float up_r = __shfl_up_sync(0xffffffffu, r[3], 1);
        float up_o = __shfl_up_sync(0xffffffffu, o[3], 1);
        if (lane == 0) {
          up_r = (MULTI && w > 0) ? edge_r[((step + 1) & 1) * 8 + w - 1] : kNeg;
          up_o = (MULTI && w > 0) ? edge_o[((step + 1) & 1) * 8 + w - 1] : o[0];
        }
        // New frame of every row.  A dead row adopts the frame of the row that
        // feeds it: the NEW frame for the three in-lane neighbours (exact, local),
        // the one-step-old frame across lanes (the shuffled residual is raw).
        on[0] = (alive[0] || first_row_lane) ? o[0] + kk[0] : up_o;
        on[1] = alive[1] ? o[1] + kk[1] : on[0];
        on[2] = alive[2] ? o[2] + kk[2] : on[1];
        on[3] = alive[3] ? o[3] + kk[3] : on[2];
        const float rc0 = r[0] - kk[0], rc1 = r[1] - kk[1], rc2 = r[2] - kk[2], rc3 = r[3] - kk[3];
        const float n0 = logadd2(up_r + (x.x + (up_o - on[0])), rc0 + y.x);
        const float n1 = logadd2(rc0 + (x.y + (on[0] - on[1])), rc1 + y.y);
        const float n2 = logadd2(rc1 + (x.z + (on[1] - on[2])), rc2 + y.z);
        const float n3 = logadd2(rc2 + (x.w + (on[2] - on[3])), rc3 + y.w);
        r[0] = n0; r[1] = n1; r[2] = n2; r[3] = n3;
        o[0] = on[0]; o[1] = on[1]; o[2] = on[2]; o[3] = on[3];
        *reinterpret_cast<float4 *>(outAr + (size_t)e * P + r0) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4 *>(outAo + (size_t)e * P + r0) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        // beta_{e-1}(s') = logadd(X[e][s'+1] + beta_e(s'+1), Y[e][s'] + beta_e(s'))
        float dn_r = __shfl_down_sync(0xffffffffu, r[0], 1);
        float dn_o = __shfl_down_sync(0xffffffffu, o[0], 1);
        if (lane == 31) {
          dn_r = (MULTI && w < W - 1) ? edge_r[((step + 1) & 1) * 8 + w + 1] : kNeg;
          dn_o = (MULTI && w < W - 1) ? edge_o[((step + 1) & 1) * 8 + w + 1] : o[3];
        }
        on[3] = (alive[3] || last_row_lane) ? o[3] + kk[3] : dn_o;
        on[2] = alive[2] ? o[2] + kk[2] : on[3];
        on[1] = alive[1] ? o[1] + kk[1] : on[2];
        on[0] = alive[0] ? o[0] + kk[0] : on[1];
        const float rc0 = r[0] - kk[0], rc1 = r[1] - kk[1], rc2 = r[2] - kk[2], rc3 = r[3] - kk[3];
        const float a0 = rc1 + (x.y + (on[1] - on[0])), a1 = rc2 + (x.z + (on[2] - on[1])),
                    a2 = rc3 + (x.w + (on[3] - on[2])), a3 = dn_r + (xn + (dn_o - on[3]));
        const float c0 = rc0 + y.x, c1 = rc1 + y.y, c2 = rc2 + y.z, c3 = rc3 + y.w;
        // operands of diagonal e-1, expressed in the frame `on`
        *reinterpret_cast<float4 *>(outBx + (size_t)(e - 1) * P + r0) = make_float4(a0, a1, a2, a3);
        *reinterpret_cast<float4 *>(outBy + (size_t)(e - 1) * P + r0) = make_float4(c0, c1, c2, c3);
        *reinterpret_cast<float4 *>(outBo + (size_t)(e - 1) * P + r0) = make_float4(on[0], on[1], on[2], on[3]);
        r[0] = logadd2(a0, c0); r[1] = logadd2(a1, c1);
        r[2] = logadd2(a2, c2); r[3] = logadd2(a3, c3);
        o[0] = on[0]; o[1] = on[1]; o[2] = on[2]; o[3] = on[3];
      }
      ++step;
      if (MULTI) {
        if (lane == (dir ? 0 : 31)) {
          edge_r[((step + 1) & 1) * 8 + w] = dir ? r[0] : r[3];
          edge_o[((step + 1) & 1) * 8 + w] = dir ? o[0] : o[3];
        }
        __syncthreads();
      }
      x = x2; y = y2; xn = xn2;
    }
    if (!MULTI) __syncwarp();
    if (tid == 0 && i + kStages < nchunk) issue(i + kStages, st);
  }
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

size_t chain_smem_bytes(const DpGeom &g, int *ch_out) {
  int CH = (g.P <= 512) ? kChunk : kChunk / 2;
  *ch_out = CH;
  return (size_t)kStages * 2 * CH * g.P * sizeof(float) + kStages * sizeof(uint64_t) + 32 * sizeof(float) + 64;
}

int launch_chain(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, bool both_directions,
                 cudaStream_t stream) {
  if (g.P > kRowsPerWarp * kMaxWarpsDp) return FRN_EUNSUPPORTED;
  int CH;
  const size_t smem = chain_smem_bytes(g, &CH);
  ChainParams cp{w.X, w.Y, w.ar, w.ao, w.bx, w.by, w.bo, boundary, g.k, g.P, g.Dn, g.S, g.T, CH};
  dim3 grid(g.B, both_directions ? 2 : 1);
  const int threads = g.P / kRowsPerLane;
  cudaError_t e;
  if (threads == 32) {
    e = cudaFuncSetAttribute(dp_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return note_cuda_error(e);
    dp_chain_kernel<false><<<grid, threads, smem, stream>>>(cp);
  } else {
    e = cudaFuncSetAttribute(dp_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return note_cuda_error(e);
    dp_chain_kernel<true><<<grid, threads, smem, stream>>>(cp);
  }
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
