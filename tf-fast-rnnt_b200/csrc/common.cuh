// Shared device/host helpers for the fast_rnnt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <stdlib.h>

#include "../../include/fast_rnnt_b200.h"

namespace frn {

// "minus infinity" inside the lattice kernels: finite, so that
// (-inf) - (-inf) never appears on the dependency chain; anything below
// kNegThresh is treated as -inf when results leave the kernels.  The byte
// pattern 0xF0F0F0F0 (-5.96e29) written by cudaMemsetAsync is also below it.
constexpr float kNeg = -1.0e30f;
constexpr float kNegThresh = -1.0e29f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxRowsDp = 1024;                // S + 1 <= 1024 (8 warps x 32 lanes x 4 rows)

// Experiment / cross-check overrides (FRN_DP_CHAIN, FRN_DP_SCAN, FRN_BAND_DENSE, FRN_SIMPLE_SIMT, ...) exist
// only in the -DFRN_DEBUG_HOOKS build (libfast_rnnt_b200_dbg.so, which the tests load to run two
// implementations of one stage against each other); the product library never looks at the environment.
#ifdef FRN_DEBUG_HOOKS
inline int debug_env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
constexpr int debug_env_int(const char *, int dflt) { return dflt; }
#endif

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

// thread-local last CUDA error for frn_last_cuda_error()
int note_cuda_error(cudaError_t e);
int check_launch();
// every launch of the library is written  count_launch(), kernel<<<...>>>(...);  the count exists in the
// -DFRN_DEBUG_HOOKS build only (frn_kernel_launches(): tests, the benchmark's kernels-per-step figure) - the
// product library keeps no process-global mutable state
#ifdef FRN_DEBUG_HOOKS
void count_launch();
#else
inline void count_launch() {}
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b) for finite a, b (kNeg stands in for -inf).
__device__ __forceinline__ float logadd2(float a, float b) {
  float mx = fmaxf(a, b), mn = fminf(a, b);
#ifdef FRN_ACCURATE_LOGADD
  return mx + log2f(1.0f + exp2f(mn - mx));
#else
  return mx + lg2_approx(1.0f + ex2_approx(mn - mx));
#endif
}

// ---- extended-range linear-domain arithmetic of the lattice kernels ----
// A lattice value is  m * 2^o : `m` a float32 mantissa kept within a few binades
// of 1 (0 = "minus infinity"), `o` an exact int32 exponent ("frame").  All
// re-scalings are by exact powers of two, so the only roundings are those of the
// multiply-adds themselves.  kNegI is the frame of a dead value / dead arc; sums
// of up to four of them stay inside int32.
constexpr int kNegI = -(1 << 28);
constexpr int kNegIThresh = -(1 << 27);
// 2^d for d <= 127; anything below 2^-126 flushes to +0
__device__ __forceinline__ float pow2i(int d) { return __int_as_float((max(d, -127) + 127) << 23); }
// floor(log2(m)) of a positive normal float; 0 for m == 0
__device__ __forceinline__ int expo_of(float m) { return m > 0.f ? ((__float_as_int(m) >> 23) - 127) : 0; }
// (m, o) -> same value with m in [1,2); dead values (m == 0) are left alone
__device__ __forceinline__ void normalise_pair(float &m, int &o) {
  const int bits = __float_as_int(m);
  const bool alive = m > 0.f;
  o = alive ? o + ((bits >> 23) - 127) : o;
  m = alive ? __int_as_float((bits & 0x007fffff) | 0x3f800000) : m;
}
// log2-domain arc score -> (mantissa in [1,2], exponent); dead arcs -> (0, kNegI)
__device__ __forceinline__ float2 encode_arc(float v_log2) {
  if (!(v_log2 > -1.0e8f)) return make_float2(0.f, __int_as_float(kNegI));   // total scores must stay above -2^27
  const float e = floorf(v_log2);
  return make_float2(ex2_approx(v_log2 - e), __int_as_float((int)e));   // argument in [0,1]: 2 ulp, like exp2f
}

// natural-log score of a lattice value {mantissa, frame}
__device__ __forceinline__ float lattice_score(float2 v) {
  if (!(v.x > 0.f)) return -INFINITY;
  return (float)(((double)log2f(v.x) + (double)__float_as_int(v.y)) * 0.6931471805599453);
}
// alpha * 2^(alpha frame + operand frame - total frame) / total mantissa: multiply by a backward-side
// operand (arc * beta(head)) to get the arc's occupation count
__device__ __forceinline__ float occupation_scale(float2 alpha, int operand_frame, int total_frame, float inv_total) {
  const int d = __float_as_int(alpha.y) + operand_frame - total_frame;
  return (alpha.x * inv_total) * pow2i(min(d, 96));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- streaming 128-bit accesses (data touched once) ----
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Delay penalty of rnnt_loss.py:316-321: ((t_end-1)/2 - t) * delay_penalty in
// float64 (TF true-divides int32 into float64), rounded to float32.
__host__ __device__ inline float delay_penalty_value(int t_end, int t, float delay_penalty) {
  double off = (double)(t_end - 1) / 2.0;
  return (float)((off - (double)t) * (double)delay_penalty);
}

// ---- lattice geometry shared by the DP kernels ----
struct DpGeom {
  int B, S, T, T1;  // T1 = T+1 regular recursion, T modified recursion
  int k;            // diagonal index d = t' + k*s'   (k = 1 regular, 0 modified)
  int P;            // padded row count: multiple of 32 * rpl
  int rpl;          // lattice rows per lane in the chain kernel (1, 2 or 4)
  int Dn;           // allocated diagonals, multiple of kChunk
};
constexpr int kChunk = 16;    // diagonals per bulk copy

inline DpGeom make_geom(int B, int S, int T, int T1) {
  DpGeom g;
  g.B = B; g.S = S; g.T = T; g.T1 = T1;
  g.k = (T1 == T) ? 0 : 1;
  g.rpl = (S + 1 <= 256) ? 1 : ((S + 1 <= 512) ? 2 : 4);
  if (const int r = debug_env_int("FRN_RPL", 0)) {    // experiment knob: rows per lane of the chain kernel
    if ((r == 1 && S + 1 <= 256) || (r == 2 && S + 1 <= 512) || r == 4) g.rpl = r;
  }
  g.P = round_up(S + 1, 32 * g.rpl);
  g.Dn = round_up(T + 1 + g.k * S, kChunk);
  return g;
}

// Workspace carve-up of one DP invocation (diagonal-major planes, [B][Dn][P]).
struct DpWorkspace {
  float4 *XY;          // arcs entering cell (d, s'): {px mantissa, px exponent, py mantissa, py exponent}
  float2 *A;           // forward scores alpha: {mantissa, frame}
  float4 *Bq;          // backward side: {px-arc operand, py-arc operand, their frame, unused}; operand = arc * beta(head)
  size_t bytes;
};
inline DpWorkspace carve_dp(void *base, const DpGeom &g) {
  DpWorkspace w;
  char *p = static_cast<char *>(base);
  const size_t cells = (size_t)g.B * g.Dn * g.P;
  w.XY = reinterpret_cast<float4 *>(p); p += round_up_sz(cells * sizeof(float4), 256);
  w.A = reinterpret_cast<float2 *>(p); p += round_up_sz(cells * sizeof(float2), 256);
  w.Bq = reinterpret_cast<float4 *>(p); p += round_up_sz(cells * sizeof(float4), 256);
  w.bytes = (size_t)(p - static_cast<char *>(base));
  return w;
}

}  // namespace frn
