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
//            CTAs.  Values are renormalised every 4 steps by the diagonal
//            maximum (offsets accumulated in double), which keeps |value| small
//            and the float32 result closer to float64 truth than the
//            reference's plain float32 p[] ("numerics").
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
struct ChainParams {
  const float *X, *Y;
  float *alpha;   // [B][Dn][P]        alpha~ (dir 0)
  float *bx, *by; // [B][Dn][P]        beta-side operands (dir 1), see finalize
  double *offA, *offB;
  const int32_t *boundary;
  int k, P, Dn, S, T, CH;  // CH: diagonals per bulk copy (divides kChunk)
};

__device__ __forceinline__ float max4(const float (&v)[4]) {
  return fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
}

template <bool MULTI>
__global__ void __launch_bounds__(256, 1) dp_chain_kernel(ChainParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x, dir = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5;
  const int P = p.P, CH = p.CH;
  const int stage_floats = 2 * CH * P;
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(ring + kStages * stage_floats);
  float *edge = reinterpret_cast<float *>(mbar + kStages);  // [2][8]
  float *wmaxs = edge + 16;                                 // [8]

  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  if (Sb < 0 || Tb < 0 || bd.x < 0 || bd.y < 0 || bd.z > p.S || bd.w > p.T) return;  // finalize reports it
  const int Db = Tb + p.k * Sb;
  const int nchunk = Db / CH + 1;
  const size_t plane = (size_t)b * p.Dn * P;
  const float *Xg = p.X + plane, *Yg = p.Y + plane;
  float *outA = p.alpha + plane, *outBx = p.bx + plane, *outBy = p.by + plane;
  double *offs = (dir ? p.offB : p.offA) + (size_t)b * p.Dn;
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
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = ((r0 + j) == (dir ? Sb : 0)) ? 0.f : kNeg;
  if (!dir) {
    *reinterpret_cast<float4 *>(outA + r0) = make_float4(v[0], v[1], v[2], v[3]);
    if (tid == 0) offs[0] = 0.0;
  }
  if (MULTI) {
    if (lane == (dir ? 0 : 31)) edge[8 + w] = dir ? v[0] : v[3];
    __syncthreads();
  }

  double A = 0.0;
  float pend = 0.f;
  int step = 0;
  for (int i = 0; i < nchunk; ++i) {
    const int st = i % kStages;
    const int ci = dir ? nchunk - 1 - i : i;
    mbar_wait(&mbar[st], (uint32_t)((i / kStages) & 1));
    const float *xs = ring + st * stage_floats, *ys = xs + CH * P;
    const int e_lo = max(ci * CH, 1), e_hi = min(ci * CH + CH - 1, Db);
    const int n = e_hi - e_lo + 1;
    float4 x, y;
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
      const double A_before = A;
      if (!dir) {
        // alpha_e(s') = logadd(alpha_{e-1}(s'-1) + X[e][s'], alpha_{e-1}(s') + Y[e][s'])
        float up = __shfl_up_sync(0xffffffffu, v[3], 1);
        if (lane == 0) up = (MULTI && w > 0) ? edge[((step + 1) & 1) * 8 + w - 1] : kNeg;
        const float n0 = logadd2(up + x.x, v[0] + y.x);
        const float n1 = logadd2(v[0] + x.y, v[1] + y.y);
        const float n2 = logadd2(v[1] + x.z, v[2] + y.z);
        const float n3 = logadd2(v[2] + x.w, v[3] + y.w);
        v[0] = n0; v[1] = n1; v[2] = n2; v[3] = n3;
      } else {
        // beta_{e-1}(s') = logadd(X[e][s'+1] + beta_e(s'+1), Y[e][s'] + beta_e(s'))
        float dn = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane == 31) dn = (MULTI && w < W - 1) ? edge[((step + 1) & 1) * 8 + w + 1] : kNeg;
        const float a0 = x.y + v[1], a1 = x.z + v[2], a2 = x.w + v[3], a3 = xn + dn;
        const float c0 = y.x + v[0], c1 = y.y + v[1], c2 = y.z + v[2], c3 = y.w + v[3];
        *reinterpret_cast<float4 *>(outBx + (size_t)(e - 1) * P + r0) = make_float4(a0, a1, a2, a3);
        *reinterpret_cast<float4 *>(outBy + (size_t)(e - 1) * P + r0) = make_float4(c0, c1, c2, c3);
        v[0] = logadd2(a0, c0); v[1] = logadd2(a1, c1);
        v[2] = logadd2(a2, c2); v[3] = logadd2(a3, c3);
      }
      // periodic renormalisation by the diagonal maximum (deferred 2 steps)
      if ((step & 3) == 0) {
        float m = warp_max(max4(v));
        if (MULTI) { if (lane == 0) wmaxs[w] = m; } else pend = m;
      } else if ((step & 3) == 2) {
        float m = pend;
        if (MULTI) {
          m = wmaxs[0];
          for (int j = 1; j < W; ++j) m = fmaxf(m, wmaxs[j]);
        }
        if (m < kNegThresh) m = 0.f;
        v[0] -= m; v[1] -= m; v[2] -= m; v[3] -= m;
        A += (double)m;
      }
      if (!dir) {
        *reinterpret_cast<float4 *>(outA + (size_t)e * P + r0) = make_float4(v[0], v[1], v[2], v[3]);
        if (tid == 0) offs[e] = A;
      } else {
        if (tid == 0) offs[e - 1] = A_before;
      }
      ++step;
      if (MULTI) {
        if (lane == (dir ? 0 : 31)) edge[((step + 1) & 1) * 8 + w] = dir ? v[0] : v[3];
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
  const float *alpha, *bx, *by;
  const double *offA, *offB;
  const int32_t *boundary;
  float *ans;               // [B]
  float *px_grad, *py_grad; // reference layout, may be null
  int S, T, T1, P, Dn, k;
};

__device__ __forceinline__ bool boundary_ok(const int4 &bd, int S, int T) {
  return bd.z - bd.x >= 0 && bd.w - bd.y >= 0 && bd.x >= 0 && bd.y >= 0 && bd.z <= S && bd.w <= T;
}

// total score of one utterance in log2 units (double) from the alpha chain
__device__ __forceinline__ double dp_total(const float *alpha, const double *offA, int b, int Dn, int P, int Db,
                                           int Sb) {
  const float a = alpha[((size_t)b * Dn + Db) * P + Sb];
  return (double)a + offA[(size_t)b * Dn + Db];
}

__global__ void dp_ans_kernel(FinalizeDenseParams p, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  if (!boundary_ok(bd, p.S, p.T)) { p.ans[b] = 0.f; return; }
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const double tot = dp_total(p.alpha, p.offA, b, p.Dn, p.P, Tb + p.k * Sb, Sb);
  p.ans[b] = (tot < (double)kNegThresh) ? -INFINITY : (float)(tot * 0.6931471805599453);
}

// One block: 32 rows (s) x 64 columns (t) of one utterance, absolute indices.
template <int K>
__global__ void __launch_bounds__(256) finalize_dense_kernel(FinalizeDenseParams p) {
  constexpr int TS = 32, TT = 64;
  constexpr int ND = TT + K * (TS - 1);  // diagonals touched
  constexpr int kPitch = 33;
  extern __shared__ float sm[];
  float *sa = sm, *sbx = sa + ND * kPitch, *sby = sbx + ND * kPitch, *sc = sby + ND * kPitch;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * TT, s0 = blockIdx.y * TS;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, s_end = bd.z, t_end = bd.w;
  const int Sb = s_end - s_begin, Tb = t_end - t_begin;
  const bool ok = boundary_ok(bd, p.S, p.T);
  const int noff = K ? 0 : 1;
  float *gx = p.px_grad + (size_t)b * p.S * p.T1;
  float *gy = p.py_grad + (size_t)b * (p.S + 1) * p.T;

  // tile in boundary-relative coordinates
  const int sp0 = s0 - s_begin, tp0 = t0 - t_begin;
  const int dlo = tp0 + K * sp0;
  const int Db = Tb + K * Sb;
  double tot = 0.0;
  bool dead = !ok;
  if (ok) {
    tot = dp_total(p.alpha, p.offA, b, p.Dn, p.P, Db, Sb);
    dead = tot < (double)kNegThresh;
  }
  // does the tile intersect the box of arcs at all?
  const bool inter = !dead && sp0 + TS > 0 && sp0 <= Sb && tp0 + TT > 0 && tp0 <= Tb;
  if (inter) {
    const size_t plane = (size_t)b * p.Dn * p.P;
    for (int i = threadIdx.x; i < ND * TS; i += blockDim.x) {
      const int dd = i / TS, ss = i - dd * TS;
      const int d = dlo + dd, sp = sp0 + ss;
      float a = kNeg, x = kNeg, y = kNeg;
      if (d >= 0 && d < Db && sp >= 0 && sp <= Sb) {  // arcs leave diagonals 0..Db-1
        const size_t idx = plane + (size_t)d * p.P + sp;
        a = p.alpha[idx]; x = p.bx[idx]; y = p.by[idx];
      }
      sa[dd * kPitch + ss] = a; sbx[dd * kPitch + ss] = x; sby[dd * kPitch + ss] = y;
    }
    for (int dd = threadIdx.x; dd < ND; dd += blockDim.x) {
      const int d = dlo + dd;
      float c = 0.f;
      if (d >= 0 && d < Db) c = (float)(p.offA[(size_t)b * p.Dn + d] + p.offB[(size_t)b * p.Dn + d] - tot);
      sc[dd] = c;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TS * TT; i += blockDim.x) {
    const int ss = i / TT, tt = i - ss * TT;
    const int s = s0 + ss, t = t0 + tt;
    if (s > p.S) continue;
    float vx = 0.f, vy = 0.f;
    if (inter) {
      const int sp = s - s_begin, tp = t - t_begin;
      if (sp >= 0 && tp >= 0 && sp <= Sb && tp <= Tb) {
        const int dd = tt + K * ss;
        const float a = sa[dd * kPitch + ss], c = sc[dd];
        // arc (s,t)->(s+1,t+noff): cu:727-746; arc (s,t)->(s,t+1): cu:747-753
        if (sp < Sb && tp + noff <= Tb) vx = ex2_approx(a + sbx[dd * kPitch + ss] + c);
        if (tp < Tb) vy = ex2_approx(a + sby[dd * kPitch + ss] + c);
      }
    }
    if (s < p.S && t < p.T1) gx[(size_t)s * p.T1 + t] = vx;
    if (t < p.T) gy[(size_t)s * p.T + t] = vy;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    p.ans[b] = !ok ? 0.f : (dead ? -INFINITY : (float)(tot * 0.6931471805599453));
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
  return (size_t)kStages * 2 * CH * g.P * sizeof(float) + kStages * sizeof(uint64_t) + 24 * sizeof(float) + 64;
}

int launch_chain(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, bool both_directions,
                 cudaStream_t stream) {
  if (g.P > kRowsPerWarp * kMaxWarpsDp) return FRN_EUNSUPPORTED;
  int CH;
  const size_t smem = chain_smem_bytes(g, &CH);
  ChainParams cp{w.X, w.Y, w.alpha, w.bx, w.by, w.offA, w.offB, boundary,
                 g.k, g.P, g.Dn, g.S, g.T, CH};
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
  FinalizeDenseParams fp{w.alpha, w.bx, w.by, w.offA, w.offB, boundary,
                         ans, px_grad, py_grad, g.S, g.T, g.T1, g.P, g.Dn, g.k};
  if (px_grad == nullptr || py_grad == nullptr) {
    dp_ans_kernel<<<(g.B + 127) / 128, 128, 0, stream>>>(fp, g.B);
    return check_launch();
  }
  dim3 grid((g.T + 1 + 63) / 64, (g.S + 1 + 31) / 32, g.B);
  if (g.k) {
    constexpr int ND = 64 + 31;
    const size_t smem = (size_t)(3 * ND * 33 + ND) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(finalize_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return note_cuda_error(e);
    finalize_dense_kernel<1><<<grid, 256, smem, stream>>>(fp);
  } else {
    constexpr int ND = 64;
    const size_t smem = (size_t)(3 * ND * 33 + ND) * sizeof(float);
    finalize_dense_kernel<0><<<grid, 256, smem, stream>>>(fp);
  }
  return check_launch();
}

}  // namespace frn
