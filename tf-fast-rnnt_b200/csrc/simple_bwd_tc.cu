// A9 on the 5th-generation tensor cores: the two batched contractions of the am / lm gradient
// (simple_bwd.cu has the algebra; TF autodiff through rnnt_loss.py:175-221 + _RNNTLossGrad, __init__.py:154-162)
//
//   AM side:  am_grad[t,c] = -g comb amp[t,c] * sum_s W[s,t] lmp[s,c]        M = t, N = c, K = s
//   LM side:  lm_grad[s,c] = -g comb lmp[s,c] * sum_t W[s,t] amp[t,c]        M = s, N = c, K = t
//
// with amp = exp(am - ammax), lmp = exp(lm - lmmax) formed in the kernel and W = G / Z from bwd_weights_kernel
// (zero-padded [B][S1p][Tp] so that no tile needs a bounds check).  One CTA = one (utterance, 128-row M tile,
// TN-column N tile); both sides run in ONE grid (the long-K LM tiles first).  Per 64-wide K slice:
//   1. all 512 threads load their 16/32-byte pieces of the W tile and of the am / lm tile straight from
//      global memory into registers (issued while the previous slice's MMAs run),
//   2. turn them into the three-term bfloat16 split h + m + l (as the forward kernel, logprobs_simple_tc.cu)
//      and store them as SWIZZLE_128B operand tiles: the exp() operand (n contiguous in memory) and the AM
//      side's W operand (m contiguous) MN-major, the LM side's W operand (k contiguous) K-major - the layouts
//      tcgen05 reads as they are, no transposition anywhere,
//   3. one thread issues 6 products x 4 k16 steps of tcgen05.mma into one 128 x TN float32 accumulator in
//      tensor memory (all slices accumulate there: the occupation-weighted sums have no cancellation, the
//      tolerance is 1e-4) and commits to an mbarrier.
// Epilogue: tcgen05.ld 32 columns at a time, transposed through a per-warp shared-memory patch so that a warp
// reads x[m][n..n+31] and writes out[m][n..n+31] as whole 128-byte lines, scaled by -g comb exp(x - xmax).
#include <cuda.h>

#include "common.cuh"
#include "launchers.h"
#include "simple_bwd_params.cuh"

namespace frn {
namespace bt {
constexpr int TM = 128;         // MMA M: frames (AM side) / lattice rows (LM side) per CTA
constexpr int KS = 64;          // K slice = one 128-byte swizzle row of bf16
constexpr int kThreads = 512;
constexpr uint32_t kOpA = TM * KS * 2;            // one split term of the W operand
constexpr uint32_t kPanel = 64 * KS * 2;          // 64 MN elements x 64 k: 8 KB (MN-major operand panel)

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// SWIZZLE_128B shared-memory matrix descriptors (cute::UMMA::SmemDescriptor, version 1, layout type 2).
// K-major: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major: 64-element MN panels `lbo` bytes apart,
// 8-k groups 1024 B apart (SBO) - canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t smem_addr, uint32_t lbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16; bit 15 / 16: A / B MN-major
__device__ __forceinline__ constexpr uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// eight float32 -> three bf16x8 (h, m, l): h = top 16 bits, m = top 16 bits of the exact remainder, l likewise
__device__ __forceinline__ void split_store(const float (&x)[8], unsigned char *dst, uint32_t term_stride) {
  uint32_t hb[8], mb[8], lb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hb[e] = __float_as_uint(x[e]) & 0xFFFF0000u;
    const float r1 = x[e] - __uint_as_float(hb[e]);
    mb[e] = __float_as_uint(r1);
    const float r2 = r1 - __uint_as_float(mb[e] & 0xFFFF0000u);
    lb[e] = __float_as_uint(r2);
  }
  auto pack = [](const uint32_t (&v)[8]) {   // upper halves: element e low, e+1 high
    return make_uint4(__byte_perm(v[0], v[1], 0x7632), __byte_perm(v[2], v[3], 0x7632),
                      __byte_perm(v[4], v[5], 0x7632), __byte_perm(v[6], v[7], 0x7632));
  };
  *reinterpret_cast<uint4 *>(dst) = pack(hb);
  *reinterpret_cast<uint4 *>(dst + term_stride) = pack(mb);
  *reinterpret_cast<uint4 *>(dst + 2 * term_stride) = pack(lb);
}
// byte offset of 16-byte chunk `j` (0..7) of row `r` inside a SWIZZLE_128B panel of 8-row atoms
__device__ __forceinline__ uint32_t swz(int r, int j) {
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((j ^ (r & 7)) << 4);
}

// NST operand stages: with two, the conversion of slice k+1 runs while the tensor core works on slice k
template <int TN, int NST> struct Smem {
  static constexpr uint32_t kOpB = TN * KS * 2;
  static constexpr uint32_t kStage = 3 * kOpA + 3 * kOpB;       // (h, m, l) x (W operand, exp operand)
  static constexpr uint32_t kOffB = 3 * kOpA, kOffSmall = NST * kStage;
  static constexpr uint32_t kBytes = kOffSmall + 64 + 1024;     // + alignment slack
};

// One tile.  AM_SIDE: M = t, K = s, x = am, y = lm.  Otherwise: M = s, K = t, x = lm, y = am.
template <bool AM_SIDE, int TN, int NST>
__device__ __forceinline__ void bwd_tc_tile(const BwdParams &p, int b, int m0, int n0, unsigned char *smem) {
  using L = Smem<TN, NST>;
  static_assert(NST == 2, "two operand stages: full[2] / done[2]");
  // Roles: warps 0..14 convert (a piece = 8 consecutive values: 32 bytes in, 3 x 16 bytes out), warp 15 issues the
  // MMAs.  No block barrier in the loop: converters -> issuer through full[stage] (one arrival per converter
  // warp), issuer -> converters through the tcgen05.commit on done[stage].
  constexpr int kConvWarps = kThreads / 32 - 1, kConvThreads = kConvWarps * 32;
  constexpr int kTotal = (TM + TN) * KS / 8;                                   // pieces of one K slice
  constexpr int kPieces = (kTotal + kConvThreads - 1) / kConvThreads;
  constexpr int kNB = TN / 8;                           // 16-byte chunks along n
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool issuer = (w == kConvWarps);
  const int S1 = p.S + 1, C = p.C, T = p.T;
  const int K = AM_SIDE ? S1 : T, M = AM_SIDE ? T : S1;
  const int nk = (K + KS - 1) / KS;
  const float *Wb = p.W + (size_t)b * p.S1p * p.Tp;
  const float *y = AM_SIDE ? p.lm + (size_t)b * S1 * C : p.am + (size_t)b * T * C;
  const float *ymax = AM_SIDE ? p.lmmax + (size_t)b * S1 : p.ammax + (size_t)b * T;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::kOffSmall), *done = full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kOffSmall + 32);

  if (tid == 0) {
    mbar_init(&full[0], kConvWarps); mbar_init(&full[1], kConvWarps);
    mbar_init(&done[0], 1); mbar_init(&done[1], 1);
    mbar_fence_init();
  }
  if (w == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  float4 rr[kPieces][2];
  float rmax[kPieces];
  auto load_slice = [&](int ks) {
    const int k0 = ks * KS;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int pi = tid + i * kConvThreads;
      rr[i][0] = rr[i][1] = z;
      rmax[i] = 0.f;
      if (pi < TM * KS / 8) {                           // W operand (zero-padded: no bounds)
        const float *src;
        if (AM_SIDE) { const int k = pi >> 4, jm = pi & 15; src = Wb + (size_t)(k0 + k) * p.Tp + m0 + jm * 8; }
        else         { const int m = pi >> 3, j = pi & 7;   src = Wb + (size_t)(m0 + m) * p.Tp + k0 + j * 8; }
        rr[i][0] = __ldg(reinterpret_cast<const float4 *>(src));
        rr[i][1] = __ldg(reinterpret_cast<const float4 *>(src) + 1);
      } else if (pi < kTotal) {                         // exp operand
        const int q = pi - TM * KS / 8, k = q / kNB, jn = q - k * kNB;
        const int kk = k0 + k, n = n0 + jn * 8;
        if (kk < K) {
          const float4 *src = reinterpret_cast<const float4 *>(y + (size_t)kk * C + n);
          if (n < C) rr[i][0] = __ldg(src);             // C % 4 == 0: a float4 is inside or outside the row
          if (n + 4 < C) rr[i][1] = __ldg(src + 1);
          rmax[i] = __ldg(ymax + kk);
        }
      }
    }
  };
  auto store_slice = [&](int ks, unsigned char *stage) {
    const int k0 = ks * KS;
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int pi = tid + i * kConvThreads;
      const float v[8] = {rr[i][0].x, rr[i][0].y, rr[i][0].z, rr[i][0].w, rr[i][1].x, rr[i][1].y, rr[i][1].z, rr[i][1].w};
      if (pi < TM * KS / 8) {
        uint32_t off;
        if (AM_SIDE) { const int k = pi >> 4, jm = pi & 15; off = (uint32_t)(jm >> 3) * kPanel + swz(k, jm & 7); }
        else         { const int m = pi >> 3, j = pi & 7;   off = swz(m, j); }
        split_store(v, stage + off, kOpA);
      } else if (pi < kTotal) {
        const int q = pi - TM * KS / 8, k = q / kNB, jn = q - k * kNB;
        const int kk = k0 + k, n = n0 + jn * 8;
        const float nmx = -rmax[i] * kLog2e;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = (kk < K && n + e < C) ? ex2_approx(fmaf(v[e], kLog2e, nmx)) : 0.f;
        split_store(x, stage + L::kOffB + (uint32_t)(jn >> 3) * kPanel + swz(k, jn & 7), L::kOpB);
      }
    }
  };

  constexpr uint32_t id = idesc(TM, TN, AM_SIDE, true);
  if (!issuer) load_slice(0);
  for (int ks = 0; ks < nk; ++ks) {
    const int st = ks & 1;
    unsigned char *stage = smem + (uint32_t)st * L::kStage;
    if (!issuer) {
      if (ks >= 2) {                             // the MMAs of slice ks - 2 have read this stage
        mbar_wait_bounded(&done[st], (uint32_t)(((ks >> 1) - 1) & 1));
        tc_fence_after();
      }
      store_slice(ks, stage);
      if (ks + 1 < nk) load_slice(ks + 1);       // registers are free again: in flight until the next conversion
      fence_async_smem();                        // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    } else {
      mbar_wait_bounded(&full[st], (uint32_t)((ks >> 1) & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(stage), b_base = smem_u32(stage + L::kOffB);
        // small products first, h*h last (as the forward kernel)
        const int ia[6] = {2, 0, 1, 1, 0, 0}, ib[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
          for (int k16 = 0; k16 < KS / 16; ++k16) {
            const uint64_t ad = AM_SIDE ? desc_mnmajor(a_base + ia[c] * kOpA + k16 * 2048, kPanel)
                                        : desc_kmajor(a_base + ia[c] * kOpA + k16 * 32);
            const uint64_t bd = desc_mnmajor(b_base + ib[c] * L::kOpB + k16 * 2048, kPanel);
            umma_bf16(tmem_d, ad, bd, id, (ks == 0 && c == 0 && k16 == 0) ? 0u : 1u);
          }
        }
        umma_commit(&done[st]);
      }
      __syncwarp();
    }
  }
  // the last commit: all MMAs have finished (commits complete in issue order)
  mbar_wait_bounded(&done[(nk - 1) & 1], (uint32_t)(((nk - 1) >> 1) & 1));
  tc_fence_after();
  __syncthreads();                               // every warp has left the operand stages: the epilogue reuses them

  // ---- epilogue ----
  const int q = w & 3, part = w >> 2;            // TMEM lane quarter, column quarter
  constexpr int kColsPerWarp = TN / 4;
  float *patch = reinterpret_cast<float *>(smem) + w * (32 * 33);    // operands are free: 16 x 4.2 KB
  const float *x = AM_SIDE ? p.am + (size_t)b * T * C : p.lm + (size_t)b * S1 * C;
  const float *xmax = AM_SIDE ? p.ammax + (size_t)b * T : p.lmmax + (size_t)b * S1;
  float *out = AM_SIDE ? p.am_grad + (size_t)b * T * C : p.lm_grad + (size_t)b * S1 * C;
  const float scale = -(p.scores_grad ? p.scores_grad[b] : 1.f) * p.comb;
  const int mrow = m0 + q * 32 + lane;
  const float my_nmx = (mrow < M) ? -__ldg(xmax + mrow) * kLog2e : 0.f;
#pragma unroll
  for (int cc = 0; cc < kColsPerWarp / 32; ++cc) {
    const int c0 = part * kColsPerWarp + cc * 32;
    if (n0 + c0 < C) {                           // warp-uniform
      float v[32];
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) patch[lane * 33 + e] = v[e];
      __syncwarp();
      const int n = n0 + c0 + lane;
      // all 32 rows' loads in flight before the first is used (the accumulator values sit in the patch, the
      // registers are free): the AM-side tiles are two K slices and 64 KB of epilogue traffic - latency is the cost
      float xv[32];
      const float *xp = x + (size_t)(m0 + q * 32) * C + n;
      float *op = out + (size_t)(m0 + q * 32) * C + n;
      const int rows_ok = (n < C) ? min(32, M - (m0 + q * 32)) : 0;       // warp-uniform in r, per-lane in n
#pragma unroll
      for (int r = 0; r < 32; ++r) xv[r] = (r < rows_ok) ? __ldg(xp + (size_t)r * C) : 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const float nmx = __shfl_sync(0xffffffffu, my_nmx, r);
        if (r < rows_ok) op[(size_t)r * C] = scale * ex2_approx(fmaf(xv[r], kLog2e, nmx)) * patch[r * 33 + lane];
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) tmem_dealloc(tmem_d, 256);
}

// grid.x = LM tiles (long K, scheduled first) followed by AM tiles
template <int TN_AM, int TN_LM>
__global__ void __launch_bounds__(kThreads, 1) bwd_contract_tc_kernel(BwdParams p, int lm_tiles, int lm_ntiles,
                                                                      int am_mtiles, int am_ntiles) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  int i = blockIdx.x;
  if (i < lm_tiles) {
    const int per_b = lm_ntiles * ((p.S + 1 + TM - 1) / TM);
    const int b = i / per_b, r = i - b * per_b;
    bwd_tc_tile<false, TN_LM, 2>(p, b, (r / lm_ntiles) * TM, (r % lm_ntiles) * TN_LM, smem);
  } else {
    i -= lm_tiles;
    const int per_b = am_mtiles * am_ntiles;
    const int b = i / per_b, r = i - b * per_b;
    bwd_tc_tile<true, TN_AM, 2>(p, b, (r / am_ntiles) * TM, (r % am_ntiles) * TN_AM, smem);
  }
}
}  // namespace bt

bool simple_bwd_tc_applicable(const BwdParams &p) {
  return p.C % 4 == 0 && ((reinterpret_cast<uintptr_t>(p.am) | reinterpret_cast<uintptr_t>(p.lm) |
                           reinterpret_cast<uintptr_t>(p.am_grad) | reinterpret_cast<uintptr_t>(p.lm_grad)) & 15u) == 0;
}

int launch_bwd_contract_tc(const BwdParams &p, cudaStream_t stream) {
  using namespace bt;
  if (!simple_bwd_tc_applicable(p)) return FRN_EUNSUPPORTED;
  constexpr int TN_AM = 128, TN_LM = 128;
  const int S1 = p.S + 1;
  const int lm_ntiles = (p.C + TN_LM - 1) / TN_LM, lm_mtiles = (S1 + TM - 1) / TM;
  const int am_ntiles = (p.C + TN_AM - 1) / TN_AM, am_mtiles = (p.T + TM - 1) / TM;
  const long long lm_tiles = (long long)p.B * lm_ntiles * lm_mtiles, am_tiles = (long long)p.B * am_ntiles * am_mtiles;
  if (lm_tiles + am_tiles > 0x7fffffffLL) return FRN_EUNSUPPORTED;
  auto kernel = bwd_contract_tc_kernel<TN_AM, TN_LM>;
  constexpr uint32_t smem = Smem<TN_AM, 2>::kBytes > Smem<TN_LM, 2>::kBytes ? Smem<TN_AM, 2>::kBytes : Smem<TN_LM, 2>::kBytes;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return note_cuda_error(e);
  count_launch(), kernel<<<(unsigned)(lm_tiles + am_tiles), kThreads, smem, stream>>>(p, (int)lm_tiles, lm_ntiles,
                                                                                    am_mtiles, am_ntiles);
  return check_launch();
}

}  // namespace frn
