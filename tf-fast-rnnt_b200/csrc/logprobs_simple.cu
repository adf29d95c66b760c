// Simple / smoothed log-probs (A1, A2): the am+lm normaliser
//   norm[b,s,t] = log( sum_c exp(lm[b,s,c]-lmmax[b,s]) * exp(am[b,t,c]-ammax[b,t]) + tiny )
//                 + lmmax[b,s] + ammax[b,t]
// fused with the max-shift, the log and the px/py symbol gather
// (rnnt_loss.py:175-221, 1266-1365: ~10 TF kernels and 3 lattice-sized
// temporaries in the reference).
//
// This translation unit holds the float32 SIMT implementation (exact float32
// products, the parity baseline).  The tcgen05/TMA split-precision version of
// the contraction lives in logprobs_simple_tc.cu when enabled; both share the
// row-statistics kernels and the epilogue below.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "launchers.h"
#include "simple_params.cuh"

namespace frn {

// tf.math.nextafter(0., 1.) in float32 (rnnt_loss.py:181)
__device__ __forceinline__ float tiny_f32() { return __int_as_float(1); }

// ---------------------------------------------------------------------------
// row statistics: max (and sum of exp(x-max)) over C, one warp per row
// ---------------------------------------------------------------------------
// Rows of lm ([rows_lm][C]) and of am ([rows_am][C]) in one launch; 128-bit streaming loads when
// the rows are 16-byte aligned (C % 4 == 0), all loads of a row issued before the reduction.
// `pxam_t` != null: the warp that streams am row (b,t) also leaves am[b,t,symbols[b,s]], s < S, in
// pxam_t[b][t][s] (the row is hot in L2 / L1 at that moment; the tensor-core kernel's epilogue reads its
// frame's S values as one contiguous run instead of gathering across rows).
// `split` != null (tensor-core normaliser): the warp also leaves its row as the two-term float16 operand the
// tensor cores contract, p = exp(x - rowmax) * 2^15 = h + l' * 2^-11 with h = fp16(p), l' = fp16((p - h) * 2^11)
// (SplitPlanes; pitch Cp halves per row): every probability is exponentiated and split ONCE here instead of once
// per 128 x 112 tile of the contraction (12 x for lm, 4 x for am at the c4 shape).
// am / lm element types: float32 (the reference's), bfloat16 and float16 (consumed as they are -
// the row-statistics kernel is the only kernel of the tensor-core path that reads am / lm at all).
template <typename T> __device__ __forceinline__ float elem_to_f(T v);
template <> __device__ __forceinline__ float elem_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float elem_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float elem_to_f<__half>(__half v) { return __half2float(v); }
// four consecutive elements (16-byte / 8-byte aligned), streamed
template <typename T> __device__ __forceinline__ float4 load4_stream(const T *p);
template <> __device__ __forceinline__ float4 load4_stream<float>(const float *p) {
  return ld_stream_f4(reinterpret_cast<const float4 *>(p));
}
template <> __device__ __forceinline__ float4 load4_stream<__nv_bfloat16>(const __nv_bfloat16 *p) {
  uint32_t a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  return make_float4(__uint_as_float(a << 16), __uint_as_float(a & 0xFFFF0000u), __uint_as_float(b << 16),
                     __uint_as_float(b & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 load4_stream<__half>(const __half *p) {
  uint32_t a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&a));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&b));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <typename T> __device__ __forceinline__ float4 load4_cached(const T *p);
template <> __device__ __forceinline__ float4 load4_cached<float>(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
template <> __device__ __forceinline__ float4 load4_cached<__nv_bfloat16>(const __nv_bfloat16 *p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 load4_cached<__half>(const __half *p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
  const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float split4(const float4 &x, float nmx, uint2 &h, uint2 &l) {
  const float p0 = ex2_approx(fmaf(x.x, kLog2e, nmx)), p1 = ex2_approx(fmaf(x.y, kLog2e, nmx));
  const float p2 = ex2_approx(fmaf(x.z, kLog2e, nmx)), p3 = ex2_approx(fmaf(x.w, kLog2e, nmx));
  const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((p0 - f01.x) * 2048.f, (p1 - f01.y) * 2048.f);
  const __half2 l23 = __floats2half2_rn((p2 - f23.x) * 2048.f, (p3 - f23.y) * 2048.f);
  h = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
  l = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
  return (p0 + p1) + (p2 + p3);                 // 2^15 * sum of the four probabilities (accuracy guard of the normaliser)
}

// `gat` (tensor-core path): the three other values the normaliser's epilogue needs of am / lm - am[b,t,blank],
// lm[b,s,blank], lm[b,s,symbols[b,s]] - so that kernel never touches am / lm itself.
template <typename TE>
__global__ void __launch_bounds__(256) rowstats_kernel(const TE *lm, int rows_lm, const TE *am, int rows_am,
                                                       int C, float *lmmax, float *lmsum, float *ammax,
                                                       const int32_t *symbols = nullptr, int S = 0, int T = 1,
                                                       float *pxam_t = nullptr, SplitPlanes split = SplitPlanes(),
                                                       RowGathers gat = RowGathers()) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows_lm + rows_am) return;
  const bool is_lm = row < rows_lm;
  if (!is_lm) row -= rows_lm;
  const TE *src = (is_lm ? lm : am) + (size_t)row * C;
  // am[b,t,sym[b,s]] for the normaliser's epilogue (SimpleParams::pxam_t): the symbol indices are fetched first,
  // the gathers are issued right behind the row's own vector loads (same sectors, one memory round trip for
  // both) and stored at the end - a gather in front of the row loads was a second, dependent round trip.
  const bool gather = !is_lm && pxam_t != nullptr;
  constexpr int kG = 4;                         // gathers in flight per lane beside the row (S <= 128: all of them)
  int gidx[kG];
  float gval[kG];
  const int32_t *sym = gather ? symbols + (size_t)(row / T) * S : nullptr;
  if (gather) {
#pragma unroll
    for (int u = 0; u < kG; ++u) {
      const int s = lane + 32 * u;
      gidx[u] = s < S ? sym[s] : -1;
    }
  }
  // the blank / symbol values of the row for the normaliser's epilogue (RowGathers): requested with the other
  // gathers, stored at the end (a load at the end of the warp's life was one more exposed memory round trip)
  int sym_c = -1;
  if (gat.am_term && is_lm) {
    const int S1 = S + 1, b = row / S1, sl = row - b * S1;
    sym_c = sl < S ? symbols[(size_t)b * S + sl] : -1;
  }
  float term_v = 0.f, sym_v = 0.f;
  auto issue_gathers = [&]() {
    if (gather) {
#pragma unroll
      for (int u = 0; u < kG; ++u) gval[u] = (gidx[u] >= 0 && gidx[u] < C) ? elem_to_f(src[gidx[u]]) : 0.f;
    }
    if (gat.am_term) {
      term_v = elem_to_f(src[gat.term]);
      sym_v = (sym_c >= 0 && sym_c < C) ? elem_to_f(src[sym_c]) : 0.f;
    }
  };
  float *rmax = is_lm ? lmmax : ammax;
  float *rsum = is_lm ? lmsum : nullptr;
  float m = -INFINITY;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & (4 * sizeof(TE) - 1)) == 0);
  if (vec && C <= 4 * 32 * 8) {
    const int nv = C / 4;
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = u * 32 + lane;
      v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c < nv) v[u] = load4_stream(src + 4 * c);
    }
    issue_gathers();
#pragma unroll
    for (int u = 0; u < 8; ++u) m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
    m = warp_max(m);
    if (rsum) {
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u * 32 + lane < nv) s += (expf(v[u].x - m) + expf(v[u].y - m)) + (expf(v[u].z - m) + expf(v[u].w - m));
      s = warp_sum(s);
      if (lane == 0) rsum[row] = s;
    }
    if (split.Cp) {
      uint2 *hrow = reinterpret_cast<uint2 *>((is_lm ? split.lmh : split.amh) + (size_t)row * split.Cp);
      uint2 *lrow = reinterpret_cast<uint2 *>((is_lm ? split.lml : split.aml) + (size_t)row * split.Cp);
      const float nmx = fmaf(-m, kLog2e, 15.f);
      float psum = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = u * 32 + lane;
        if (c < nv) {
          uint2 h, l;
          psum += split4(v[u], nmx, h, l);
          hrow[c] = h; lrow[c] = l;
        }
      }
      psum = warp_sum(psum);
      if (lane == 0 && gat.am_sum) (is_lm ? gat.lm_sum : gat.am_sum)[row] = psum * 0x1p-15f;
    }
  } else if (vec) {
    // long rows (large vocabularies): three passes of 128-bit loads over a row that stays in L1 / L2
    issue_gathers();
    const int nv = C / 4;
    for (int c = lane; c < nv; c += 32) {
      const float4 x = load4_cached(src + 4 * c);
      m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
    }
    m = warp_max(m);
    if (rsum) {
      float s = 0.f;
      for (int c = lane; c < nv; c += 32) {
        const float4 x = load4_cached(src + 4 * c);
        s += (expf(x.x - m) + expf(x.y - m)) + (expf(x.z - m) + expf(x.w - m));
      }
      s = warp_sum(s);
      if (lane == 0) rsum[row] = s;
    }
    if (split.Cp) {
      uint2 *hrow = reinterpret_cast<uint2 *>((is_lm ? split.lmh : split.amh) + (size_t)row * split.Cp);
      uint2 *lrow = reinterpret_cast<uint2 *>((is_lm ? split.lml : split.aml) + (size_t)row * split.Cp);
      const float nmx = fmaf(-m, kLog2e, 15.f);
      float psum = 0.f;
      for (int c = lane; c < nv; c += 32) {
        uint2 h, l;
        psum += split4(load4_cached(src + 4 * c), nmx, h, l);
        hrow[c] = h; lrow[c] = l;
      }
      psum = warp_sum(psum);
      if (lane == 0 && gat.am_sum) (is_lm ? gat.lm_sum : gat.am_sum)[row] = psum * 0x1p-15f;
    }
  } else {
    issue_gathers();
    for (int c = lane; c < C; c += 32) m = fmaxf(m, elem_to_f(src[c]));
    m = warp_max(m);
    if (rsum) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += expf(elem_to_f(src[c]) - m);
      s = warp_sum(s);
      if (lane == 0) rsum[row] = s;
    }
  }
  if (lane == 0) rmax[row] = m;
  if (gather) {
    float *dst = pxam_t + (size_t)row * S;
#pragma unroll
    for (int u = 0; u < kG; ++u)
      if (lane + 32 * u < S) dst[lane + 32 * u] = gval[u];
    for (int s = lane + 32 * kG; s < S; s += 32) {      // long label sequences: the row is in L1 / L2 by now
      const int c = sym[s];
      dst[s] = (c >= 0 && c < C) ? elem_to_f(src[c]) : 0.f;
    }
  }
  if (gat.am_term && lane == 0) {
    if (is_lm) {
      gat.lm_term[row] = term_v;
      gat.lm_sym[row] = sym_v;
    } else {
      gat.am_term[row] = term_v;
    }
  }
}

// unigram[c] = mean_rows( exp(lm-lmmax)/lmsum ) + tiny  (rnnt_loss.py:1279-1280),
// deterministic column reduction: block = 32 columns x 8 row-striding warps.
// Batch sharded by utterance: the mean runs over the GLOBAL batch, so the column sums and the row
// count are exposed (`sums_out` [C+1], frn_smoothed_unigram_sums), all-reduced by the caller and handed back
// (`ext_sums`): then unigram = ext_sums[c] / ext_sums[C] + tiny and no row is read here.
template <typename TE>
__global__ void __launch_bounds__(256) unigram_kernel(const TE *lm, const float *lmmax, const float *lmsum,
                                                      int rows, int C, const float *ext_sums, float *sums_out,
                                                      float *unigram, float *log_unigram) {
  __shared__ float part[8][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < C && !ext_sums)
    for (int r = w; r < rows; r += 8) acc += expf(elem_to_f(lm[(size_t)r * C + c]) - lmmax[r]) / lmsum[r];
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][lane];
    if (sums_out) {
      sums_out[c] = s;
      if (c == 0) sums_out[C] = (float)rows;
      return;
    }
    const float u = (ext_sums ? ext_sums[c] / ext_sums[C] : s / (float)rows) + tiny_f32();
    unigram[c] = u;
    log_unigram[c] = logf(u);
  }
}

// amonly[b,t] = log( sum_c exp(am-ammax) * unigram[c] ) + ammax  (rnnt_loss.py:1281-1286)
template <typename TE>
__global__ void __launch_bounds__(256) amonly_kernel(const TE *am, const float *ammax, const float *unigram,
                                                     int rows, int C, float *amonly) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TE *src = am + (size_t)row * C;
  const float m = ammax[row];
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(elem_to_f(src[c]) - m) * unigram[c];
  s = warp_sum(s);
  if (lane == 0) amonly[row] = logf(s) + m;
}

// ---------------------------------------------------------------------------
// contraction + epilogue.  Block: 64 (s) x 64 (t) tile of one utterance,
// 256 threads, 4x4 micro-tile per thread, K chunks of 32.
// ---------------------------------------------------------------------------
constexpr int kTile = 64, kBK = 32, kPitchG = 68;

__global__ void __launch_bounds__(256) simple_logprobs_kernel(SimpleParams p) {
  __shared__ __align__(16) float As[kBK * kPitchG];  // lm probs, [k][s]
  __shared__ __align__(16) float Bs[kBK * kPitchG];  // am probs, [k][t]
  __shared__ float s_lmmax[kTile], s_ammax[kTile];
  const int b = blockIdx.z, s0 = blockIdx.y * kTile, t0 = blockIdx.x * kTile;
  const int tid = threadIdx.x;
  const int S1 = p.S + 1, C = p.C;
  const float *lmb = p.lm + (size_t)b * S1 * C;
  const float *amb = p.am + (size_t)b * p.T * C;
  if (tid < kTile) {
    const int s = s0 + tid, t = t0 + tid;
    s_lmmax[tid] = (s < S1) ? p.lmmax[(size_t)b * S1 + s] : 0.f;
    s_ammax[tid] = (t < p.T) ? p.ammax[(size_t)b * p.T + t] : 0.f;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 3, lk = (tid & 7) * 4;  // loader: row (0..31), k offset
  for (int k0 = 0; k0 < C; k0 += kBK) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = lr + half * 32;
      const int s = s0 + r, t = t0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = k0 + lk + j;
        float a = 0.f, bb = 0.f;
        if (c < C) {
          if (s < S1) a = expf(lmb[(size_t)s * C + c] - s_lmmax[r]);
          if (t < p.T) bb = expf(amb[(size_t)t * C + c] - s_ammax[r]);
        }
        As[(lk + j) * kPitchG + r] = a;
        Bs[(lk + j) * kPitchG + r] = bb;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[k * kPitchG + ty * 4]);
      const float4 c = *reinterpret_cast<const float4 *>(&Bs[k * kPitchG + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue: log, un-shift, symbol / blank gather (rnnt_loss.py:186-216) ----
  const int t_end = p.boundary[4 * b + 3];
  const int32_t *symb = p.symbols + (size_t)b * p.S;
  float *pxb = p.px + (size_t)b * p.S * p.T1;
  float *pyb = p.py + (size_t)b * S1 * p.T;
  float py_am[4], amonly[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int t = t0 + tx * 4 + j;
    py_am[j] = (t < p.T) ? amb[(size_t)t * C + p.term] : 0.f;
    amonly[j] = (p.smoothed && t < p.T) ? p.amonly[(size_t)b * p.T + t] : 0.f;
  }
  const float logu_term = p.smoothed ? p.logu[p.term] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = s0 + ty * 4 + i;
    if (s >= S1) continue;
    const float lmmax = s_lmmax[ty * 4 + i];
    const float py_lm = lmb[(size_t)s * C + p.term];
    const int sym = (s < p.S) ? symb[s] : 0;
    const float px_lm = (s < p.S) ? lmb[(size_t)s * C + sym] : 0.f;
    float lmonly = 0.f, logu_sym = 0.f;
    if (p.smoothed) {
      lmonly = logf(p.lmsum[(size_t)b * S1 + s]) + lmmax;
      logu_sym = p.logu[sym];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t < p.T) {
        const float norm = logf(acc[i][j] + tiny_f32()) + lmmax + s_ammax[tx * 4 + j];
        float py = py_am[j] + py_lm - norm;
        if (p.smoothed)
          py = py * p.comb + (py_lm - lmonly) * p.lm_scale + (py_am[j] + logu_term - amonly[j]) * p.am_scale;
        pyb[(size_t)s * p.T + t] = py;
        if (s < p.S) {
          const float px_am = amb[(size_t)t * C + sym];
          float px = px_am + px_lm - norm;
          if (p.smoothed)
            px = px * p.comb + (px_lm - lmonly) * p.lm_scale + (px_am + logu_sym - amonly[j]) * p.am_scale;
          if (p.rnnt_type == FRN_REGULAR && t == t_end) px = -INFINITY;  // fix_for_boundary, :28-61
          pxb[(size_t)s * p.T1 + t] = px;
        }
      } else if (t == p.T && p.T1 == p.T + 1 && s < p.S) {
        pxb[(size_t)s * p.T1 + t] = -INFINITY;  // regular: one-past-the-last frame, :193-203
      }
    }
  }
}

// constrained: px[b,s,t] += py[b,s+1,t]  (rnnt_loss.py:221, 1365)
__global__ void __launch_bounds__(256) constrained_fix_kernel(float *px, const float *py, int B, int S, int T) {
  const size_t n = (size_t)B * S * T;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / ((size_t)S * T));
  const size_t rem = i - (size_t)b * S * T;
  px[i] += py[(size_t)b * (S + 1) * T + T + rem];
}

// Long rows (1024 < C <= 8192, large vocabularies): one BLOCK per row, the row held in registers (up to 8 float4
// per thread), so the row is read from memory once for the maximum, the sum and the float16 split.  With a warp
// per row the three passes re-read 20 KB rows that ~10 k resident warps had long pushed out of L2 (the kernel
// read am and lm twice from DRAM at the c4 shape).
template <typename TE>
__global__ void __launch_bounds__(256) rowstats_long_kernel(const TE *lm, int rows_lm, const TE *am, int rows_am,
                                                            int C, float *lmmax, float *lmsum, float *ammax,
                                                            const int32_t *symbols, int S, int T, float *pxam_t,
                                                            SplitPlanes split, RowGathers gat) {
  __shared__ float red[8];
  int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool is_lm = row < rows_lm;
  if (!is_lm) row -= rows_lm;
  const TE *src = (is_lm ? lm : am) + (size_t)row * C;
  const int nv = C / 4;
  float4 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = u * 256 + tid;
    v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c < nv) v[u] = load4_stream(src + 4 * c);
  }
  // am[b,t,sym[b,s]] (see rowstats_kernel): issued behind the row's loads
  const bool gather = !is_lm && pxam_t != nullptr;
  if (gather) {
    const int32_t *sym = symbols + (size_t)(row / T) * S;
    float *dst = pxam_t + (size_t)row * S;
    for (int s = tid; s < S; s += 256) {
      const int c = sym[s];
      dst[s] = (c >= 0 && c < C) ? elem_to_f(src[c]) : 0.f;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < 8; ++u) m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
  m = warp_max(m);
  if (lane == 0) red[w] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  float *rsum = is_lm ? lmsum : nullptr;
  if (rsum) {                                   // block-uniform
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (u * 256 + tid < nv) s += (expf(v[u].x - m) + expf(v[u].y - m)) + (expf(v[u].z - m) + expf(v[u].w - m));
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[w] = s;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i];
      rsum[row] = t;
    }
  }
  if (tid == 0) (is_lm ? lmmax : ammax)[row] = m;
  if (gat.am_term && tid == 32) {
    if (is_lm) {
      const int S1 = S + 1, b = row / S1, sl = row - b * S1;
      gat.lm_term[row] = elem_to_f(src[gat.term]);
      const int c = sl < S ? symbols[(size_t)b * S + sl] : -1;
      gat.lm_sym[row] = (c >= 0 && c < C) ? elem_to_f(src[c]) : 0.f;
    } else {
      gat.am_term[row] = elem_to_f(src[gat.term]);
    }
  }
  if (split.Cp) {
    uint2 *hrow = reinterpret_cast<uint2 *>((is_lm ? split.lmh : split.amh) + (size_t)row * split.Cp);
    uint2 *lrow = reinterpret_cast<uint2 *>((is_lm ? split.lml : split.aml) + (size_t)row * split.Cp);
    const float nmx = fmaf(-m, kLog2e, 15.f);
    float psum = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = u * 256 + tid;
      if (c < nv) {
        uint2 h, l;
        psum += split4(v[u], nmx, h, l);
        hrow[c] = h; lrow[c] = l;
      }
    }
    if (gat.am_sum) {                            // block-uniform
      psum = warp_sum(psum);
      __syncthreads();
      if (lane == 0) red[w] = psum;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        (is_lm ? gat.lm_sum : gat.am_sum)[row] = t * 0x1p-15f;
      }
    }
  }
}

// row statistics of lm ([rows_lm][C]) and am ([rows_am][C]) in one launch: a warp per row, or a block per row for
// long aligned rows
template <typename TE>
static void launch_rowstats(const TE *lm, int rows_lm, const TE *am, int rows_am, int C, float *lmmax,
                            float *lmsum, float *ammax, const int32_t *symbols, int S, int T, float *pxam_t,
                            const SplitPlanes &split, const RowGathers &gat, cudaStream_t stream) {
  const uintptr_t mask = 4 * sizeof(TE) - 1;
  const bool aligned = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm)) & mask) == 0;
  if (aligned && C > 4 * 32 * 8 && C <= 4 * 256 * 8 && rows_lm + rows_am > 0) {
    count_launch(), rowstats_long_kernel<TE><<<rows_lm + rows_am, 256, 0, stream>>>(lm, rows_lm, am, rows_am, C, lmmax,
                                                                                     lmsum, ammax, symbols, S, T, pxam_t,
                                                                                     split, gat);
  } else {
    count_launch(), rowstats_kernel<TE><<<(rows_lm + rows_am + 7) / 8, 256, 0, stream>>>(lm, rows_lm, am, rows_am, C,
                                                                                          lmmax, lmsum, ammax, symbols, S,
                                                                                          T, pxam_t, split, gat);
  }
}

// ---------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------
size_t simple_stats_bytes(int B, int S, int T, int C) {
  size_t n = 2 * round_up_sz((size_t)B * (S + 1) * sizeof(float), 256) +
             2 * round_up_sz((size_t)B * T * sizeof(float), 256) + 2 * round_up_sz((size_t)C * sizeof(float), 256) +
             round_up_sz((size_t)B * T * S * sizeof(float), 256);        // pxam_t
  // two-term float16 operands of the tensor-core contraction (SplitPlanes): h and l planes of am and of lm
  const size_t Cp = (size_t)round_up(C, 8);
  n += 2 * round_up_sz((size_t)B * T * Cp * 2, 256) + 2 * round_up_sz((size_t)B * (S + 1) * Cp * 2, 256);
  // am[b,t,blank], lm[b,s,blank], lm[b,s,symbol], and the row sums of exp(x - max) (RowGathers)
  n += 2 * round_up_sz((size_t)B * T * sizeof(float), 256) + 3 * round_up_sz((size_t)B * (S + 1) * sizeof(float), 256);
  return n;
}

// The statistics block of the smoothed log-probs on its own (used again by the backward pass):
// lmmax, lmsum, ammax, amonly, unigram, log unigram.
int launch_smoothing_stats(const float *lm, const float *am, int B, int S, int T, int C, void *stats_ws,
                           cudaStream_t stream, const float *unigram_sums) {
  const int S1 = S + 1;
  char *w = static_cast<char *>(stats_ws);
  float *lmmax = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *lmsum = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *ammax = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
  float *amonly = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
  float *unigram = reinterpret_cast<float *>(w); w += round_up_sz((size_t)C * sizeof(float), 256);
  float *logu = reinterpret_cast<float *>(w);
  launch_rowstats(lm, B * S1, am, B * T, C, lmmax, lmsum, ammax, nullptr, 0, 1, nullptr, SplitPlanes(), RowGathers(), stream);
  count_launch(), unigram_kernel<<<(C + 31) / 32, 256, 0, stream>>>(lm, lmmax, lmsum, B * S1, C, unigram_sums, nullptr,
                                                                 unigram, logu);
  count_launch(), amonly_kernel<<<(B * T + 7) / 8, 256, 0, stream>>>(am, ammax, unigram, B * T, C, amonly);
  return check_launch();
}

// sums[c] = sum over the B (S+1) lm rows of softmax(lm row)[c], sums[C] = B (S+1): this rank's share of the
// batch-global unigram of rnnt_loss.py:1279-1280
int launch_unigram_sums(const float *lm, int B, int S, int C, void *stats_ws, float *sums, cudaStream_t stream) {
  const int S1 = S + 1;
  char *w = static_cast<char *>(stats_ws);
  float *lmmax = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *lmsum = reinterpret_cast<float *>(w);
  launch_rowstats(lm, B * S1, lm, 0, C, lmmax, lmsum, lmmax, nullptr, 0, 1, nullptr, SplitPlanes(), RowGathers(), stream);
  count_launch(), unigram_kernel<<<(C + 31) / 32, 256, 0, stream>>>(lm, lmmax, lmsum, B * S1, C, nullptr, sums, nullptr,
                                                                 nullptr);
  return check_launch();
}

// One launcher for both outputs of the contraction: px/py in the reference layout (arcs == nullptr), or
// the arcs of the dense-lattice recursion written straight into its diagonal-major plane (arcs != nullptr:
// tensor-core kernel only, regular / modified; check simple_arc_plane_supported() first).
bool simple_arc_plane_supported(const void *lm, const void *am, int C, int rnnt_type) {
  if (debug_env_int("FRN_SIMPLE_SIMT", 0) == 1 || debug_env_int("FRN_SIMPLE_PXPY", 0) == 1) return false;
  return rnnt_type != FRN_CONSTRAINED && simple_logprobs_tc_applicable(lm, am, C);
}

template <typename TE>
static int launch_simple_logprobs_t(const TE *lm, const TE *am, const int32_t *symbols, const int32_t *boundary,
                                    int B, int S, int T, int C, int term, int rnnt_type, int smoothed,
                                    float lm_only_scale, float am_only_scale, float *px, float *py, void *stats_ws,
                                    cudaStream_t stream, const ArcPlaneOut *arcs, const float *unigram_sums) {
  constexpr bool kF32 = std::is_same<TE, float>::value;
  const int S1 = S + 1;
  char *w = static_cast<char *>(stats_ws);
  float *lmmax = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *lmsum = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
  float *ammax = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
  float *amonly = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
  float *unigram = reinterpret_cast<float *>(w); w += round_up_sz((size_t)C * sizeof(float), 256);
  float *logu = reinterpret_cast<float *>(w); w += round_up_sz((size_t)C * sizeof(float), 256);
  float *pxam_t = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * S * sizeof(float), 256);
  const bool tc = simple_logprobs_tc_applicable(lm, am, C) && debug_env_int("FRN_SIMPLE_SIMT", 0) != 1;
  if (!tc && !kF32) return FRN_EUNSUPPORTED;      // bf16 / fp16 am, lm: tensor-core path only (the caller widens them)
  SplitPlanes split;
  RowGathers gat;
  if (tc) {
    split.Cp = round_up(C, 8);
    const size_t am_plane = round_up_sz((size_t)B * T * split.Cp * 2, 256), lm_plane = round_up_sz((size_t)B * S1 * split.Cp * 2, 256);
    split.amh = reinterpret_cast<unsigned short *>(w); w += am_plane;
    split.aml = reinterpret_cast<unsigned short *>(w); w += am_plane;
    split.lmh = reinterpret_cast<unsigned short *>(w); w += lm_plane;
    split.lml = reinterpret_cast<unsigned short *>(w); w += lm_plane;
    gat.am_term = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
    gat.lm_term = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
    gat.lm_sym = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * S1 * sizeof(float), 256);
    gat.am_sum = reinterpret_cast<float *>(w); w += round_up_sz((size_t)B * T * sizeof(float), 256);
    gat.lm_sum = reinterpret_cast<float *>(w);
    gat.term = term;
  }
  launch_rowstats(lm, B * S1, am, B * T, C, lmmax, smoothed ? lmsum : nullptr, ammax, symbols, S, T,
                  tc ? pxam_t : nullptr, split, gat, stream);
  if (smoothed) {
    count_launch(), unigram_kernel<<<(C + 31) / 32, 256, 0, stream>>>(lm, lmmax, lmsum, B * S1, C, unigram_sums, nullptr,
                                                                   unigram, logu);
    count_launch(), amonly_kernel<<<(B * T + 7) / 8, 256, 0, stream>>>(am, ammax, unigram, B * T, C, amonly);
  }
  int rc = check_launch();
  if (rc) return rc;
  SimpleParams sp;
  sp.lm = kF32 ? reinterpret_cast<const float *>(lm) : nullptr;      // read by the SIMT kernel only
  sp.am = kF32 ? reinterpret_cast<const float *>(am) : nullptr;
  sp.gat = gat;
  sp.am_raw = am; sp.lm_raw = lm;
  sp.raw_dtype = kF32 ? FRN_F32 : (std::is_same<TE, __nv_bfloat16>::value ? FRN_BF16 : FRN_F16);
  sp.symbols = symbols; sp.boundary = boundary;
  sp.lmmax = lmmax; sp.ammax = ammax; sp.lmsum = lmsum; sp.amonly = amonly; sp.logu = logu;
  sp.px = px; sp.py = py; sp.pxam_t = tc ? pxam_t : nullptr; sp.split = split;
  sp.B = B; sp.S = S; sp.T = T; sp.T1 = (rnnt_type == FRN_REGULAR) ? T + 1 : T; sp.C = C; sp.term = term;
  sp.rnnt_type = rnnt_type; sp.smoothed = smoothed;
  // Python-float arithmetic of rnnt_loss.py:1342-1349, then cast to float32
  const double lms = (double)lm_only_scale, ams = (double)am_only_scale;
  sp.comb = (float)(1.0 - lms - ams);
  sp.lm_scale = (float)(lms == 0.0 ? 1.0e-20 : lms);
  sp.am_scale = (float)(ams == 0.0 ? 1.0e-20 : ams);
  if (arcs) {
    sp.XY = arcs->XY; sp.P = arcs->P; sp.Dn = arcs->Dn; sp.k = arcs->k; sp.delay_penalty = arcs->delay_penalty;
    return launch_simple_logprobs_tc(sp, stream);
  }
  // tensor-core path (tcgen05 + TMA); the exact-FP32 SIMT kernel serves shapes TMA cannot address (C % 4 != 0)
  rc = debug_env_int("FRN_SIMPLE_SIMT", 0) == 1 ? FRN_EUNSUPPORTED : launch_simple_logprobs_tc(sp, stream);
  if (rc == FRN_EUNSUPPORTED && kF32) {
    dim3 grid((sp.T1 + kTile - 1) / kTile, (S1 + kTile - 1) / kTile, B);
    count_launch(), simple_logprobs_kernel<<<grid, 256, 0, stream>>>(sp);
    rc = check_launch();
  }
  if (rc) return rc;
  if (rnnt_type == FRN_CONSTRAINED) {
    const size_t n = (size_t)B * S * T;
    count_launch(), constrained_fix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(px, py, B, S, T);
    rc = check_launch();
  }
  return rc;
}

int launch_simple_logprobs(const float *lm, const float *am, const int32_t *symbols, const int32_t *boundary,
                           int B, int S, int T, int C, int term, int rnnt_type, int smoothed,
                           float lm_only_scale, float am_only_scale, float *px, float *py, void *stats_ws,
                           cudaStream_t stream, const ArcPlaneOut *arcs, const float *unigram_sums) {
  return launch_simple_logprobs_t<float>(lm, am, symbols, boundary, B, S, T, C, term, rnnt_type, smoothed, lm_only_scale,
                                         am_only_scale, px, py, stats_ws, stream, arcs, unigram_sums);
}

// am / lm of any supported element type (frn_dtype); float32 takes the path above
int launch_simple_logprobs_any(const void *lm, const void *am, int dtype, const int32_t *symbols,
                               const int32_t *boundary, int B, int S, int T, int C, int term, int rnnt_type,
                               int smoothed, float lm_only_scale, float am_only_scale, float *px, float *py,
                               void *stats_ws, cudaStream_t stream, const ArcPlaneOut *arcs, const float *unigram_sums) {
  if (dtype == FRN_F32)
    return launch_simple_logprobs(static_cast<const float *>(lm), static_cast<const float *>(am), symbols, boundary, B, S,
                                  T, C, term, rnnt_type, smoothed, lm_only_scale, am_only_scale, px, py, stats_ws, stream,
                                  arcs, unigram_sums);
  if (dtype == FRN_BF16)
    return launch_simple_logprobs_t<__nv_bfloat16>(static_cast<const __nv_bfloat16 *>(lm),
                                                   static_cast<const __nv_bfloat16 *>(am), symbols, boundary, B, S, T, C,
                                                   term, rnnt_type, smoothed, lm_only_scale, am_only_scale, px, py,
                                                   stats_ws, stream, arcs, unigram_sums);
  if (dtype == FRN_F16)
    return launch_simple_logprobs_t<__half>(static_cast<const __half *>(lm), static_cast<const __half *>(am), symbols,
                                            boundary, B, S, T, C, term, rnnt_type, smoothed, lm_only_scale, am_only_scale,
                                            px, py, stats_ws, stream, arcs, unigram_sums);
  return FRN_EINVAL;
}

}  // namespace frn
