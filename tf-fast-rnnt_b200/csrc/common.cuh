// Shared device/host helpers for the fast_rnnt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>

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

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

// thread-local last CUDA error for frn_last_cuda_error()
int note_cuda_error(cudaError_t e);
int check_launch();

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
  g.P = round_up(S + 1, 32 * g.rpl);
  g.Dn = round_up(T + 1 + g.k * S, kChunk);
  return g;
}

// Workspace carve-up of one DP invocation.
struct DpWorkspace {
  float *X, *Y;        // [B][Dn][P] skewed log2-domain arc scores
  // every lattice value is held as (exact integer offset o) + (small float32 residual r)
  float *ar, *ao;      // [B][Dn][P] forward scores: residual, offset
  float *bx, *by;      // [B][Dn][P] backward-side operands (arc score + beta of the arc's head), residuals
  float *bo;           // [B][Dn][P] offset of the frame bx/by are expressed in
  size_t bytes;
};
inline DpWorkspace carve_dp(void *base, const DpGeom &g) {
  DpWorkspace w;
  char *p = static_cast<char *>(base);
  size_t plane = round_up_sz((size_t)g.B * g.Dn * g.P * sizeof(float), 256);
  w.X = reinterpret_cast<float *>(p); p += plane;
  w.Y = reinterpret_cast<float *>(p); p += plane;
  w.ar = reinterpret_cast<float *>(p); p += plane;
  w.ao = reinterpret_cast<float *>(p); p += plane;
  w.bx = reinterpret_cast<float *>(p); p += plane;
  w.by = reinterpret_cast<float *>(p); p += plane;
  w.bo = reinterpret_cast<float *>(p); p += plane;
  w.bytes = (size_t)(p - static_cast<char *>(base));
  return w;
}

}  // namespace frn
