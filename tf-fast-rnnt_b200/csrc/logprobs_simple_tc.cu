// Simple / smoothed log-probs (A1, A2) on the 5th-generation tensor cores.
//
//   norm[b,s,t] = log( sum_c exp(lm[b,s,c]-lmmax[b,s]) * exp(am[b,t,c]-ammax[b,t]) + tiny ) + maxes
//
// One CTA = one (utterance, 128-frame tile, <=112-symbol tile).  Per 64-wide
// slice of the vocabulary axis:
//   1. TMA (cp.async.bulk.tensor.3d, tensor maps over am [B][T][C] and
//      lm [B][S+1][C]) drops the raw float32 tiles into a 2-stage shared-memory
//      ring behind mbarriers;
//   2. all 512 threads turn them into probabilities exp(x - rowmax) and split
//      each into three bfloat16 terms h+m+l (24 mantissa bits), written in the
//      K-major SWIZZLE_128B layout tcgen05 reads;
//   3. one thread issues 6 tcgen05.mma (hh, hm, mh, mm, hl, lh — everything down
//      to 2^-24 relative) per 16-wide k step into a 128 x N float32 accumulator
//      in tensor memory and commits to an mbarrier.
// Epilogue: tcgen05.ld the accumulator (one lattice frame per thread), log,
// un-shift, symbol / blank gather, smoothing terms, boundary fix-ups, and store
// px/py coalesced along t in the reference layout (rnnt_loss.py:186-221,1290-1365).
//
// float32-accurate by construction (the occupation counts downstream need ~2^-21
// on the normaliser, which rules out plain bf16/tf32: DESIGN.md "numerics").
#include <cuda.h>
#ifdef FRN_TC_TIMING
#include <cstdio>
#define TCT(i) do { if (tid == 0 && blockIdx.x == 1 && blockIdx.z == 3) tct[i] = clock64(); } while (0)
#else
#define TCT(i) do { } while (0)
#endif

#include "common.cuh"
#include "simple_params.cuh"

namespace frn {

namespace tc {
constexpr int TM = 128;   // frames per CTA  (MMA M)
constexpr int TN = 112;   // symbols per CTA (MMA N, multiple of 16)
constexpr int KC = 64;    // vocabulary slice per stage = one 128-byte swizzle row of bf16
constexpr int kThreads = 512;   // 16 warps: 4 per TMEM lane quarter, 28 symbol columns each
constexpr int kTmemCols = 256;  // [0,112) accumulator, [128,240) px_am staging
constexpr int kPxCol = 128;
constexpr uint32_t kRawAmBytes = TM * KC * 4, kRawLmBytes = TN * KC * 4;
constexpr uint32_t kOpABytes = TM * KC * 2, kOpBBytes = TN * KC * 2;
// shared memory map (byte offsets from a 1024-aligned base)
constexpr uint32_t kOffA = 0;                                  // 3 x A operand (h, m, l)
constexpr uint32_t kOffB = kOffA + 3 * kOpABytes;              // 3 x B operand
constexpr uint32_t kOffRaw = kOffB + 3 * kOpBBytes;            // 2 x (am raw, lm raw)
constexpr uint32_t kRawStage = kRawAmBytes + kRawLmBytes;
constexpr uint32_t kOffSmall = kOffRaw + 2 * kRawStage;
constexpr uint32_t kSmallBytes = 5120;
constexpr uint32_t kSmemBytes = kOffSmall + kSmallBytes + 1024;  // + alignment slack

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(1024 B between 8-row groups)<<32 |
// version 1 <<46 | layout SWIZZLE_128B (2) <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// log(z + tiny) for z >= 0: exponent taken exactly, lg2.approx only sees the mantissa in
// [1,2) where its absolute error is 2^-22.6 (better than 1 ulp of the result here);
// z == 0 gives log(tiny) like the reference's log(0 + nextafter(0,1)) (rnnt_loss.py:181).
__device__ __forceinline__ float log_plus_tiny(float z) {
  const uint32_t u = __float_as_uint(z);
  const float e = (float)((int)(u >> 23) - 127);
  const float m = __uint_as_float((u & 0x007FFFFFu) | 0x3F800000u);
  const float r = (e + lg2_approx(m)) * kLn2;
  return (z < 1.1754944e-38f) ? -103.27893f : r;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// bounded mbarrier wait: a broken pipeline traps instead of hanging the GPU
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

}  // namespace tc

__global__ void __launch_bounds__(tc::kThreads, 1)
simple_logprobs_tc_kernel(const __grid_constant__ CUtensorMap map_am, const __grid_constant__ CUtensorMap map_lm,
                          SimpleParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B operands need a 1024-byte aligned base; keep the arithmetic on the
  // __shared__ array so that accesses stay LDS/STS (a pointer rebuilt from an integer
  // degrades them to generic LD/ST)
  unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#ifdef FRN_TC_TIMING
  long long tct[48];
  for (int i = 0; i < 48; ++i) tct[i] = 0;
#endif
  TCT(0);
  const int b = blockIdx.z, t0 = blockIdx.x * TM, s0 = blockIdx.y * TN;
  const int S1 = p.S + 1, C = p.C;
  const int nk = (C + KC - 1) / KC;
  const int n_rows = min(TN, round_up(S1 - s0, 16));   // MMA N for this tile (multiple of 16)

  float *s_amneg = reinterpret_cast<float *>(smem + kOffSmall);          // [128] -ammax*log2e (-inf: row masked)
  float *s_lmneg = s_amneg + TM;                                         // [112] -lmmax*log2e
  float *s_ammax = s_lmneg + TN;                                         // [128]
  float *s_lmmax = s_ammax + TM;                                         // [112]
  float *s_pxlm = s_lmmax + TN, *s_pylm = s_pxlm + TN, *s_lmonly = s_pylm + TN, *s_logusym = s_lmonly + TN;
  int *s_sym = reinterpret_cast<int *>(s_logusym + TN);                  // [112]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_sym + TN);             // raw_full[2], mma_done
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 4);

  // The raw tiles do not depend on the row statistics: the first two slices are requested before
  // anything else so that their (cold) latency hides behind the set-up loads below.
  auto issue_tma = [&](int k, int stage) {
    unsigned char *raw = smem + kOffRaw + stage * kRawStage;
    mbar_arrive_expect_tx(&bars[stage], kRawStage);
    tma_load_3d(raw, &map_am, &bars[stage], k * KC, t0, b);
    tma_load_3d(raw + kRawAmBytes, &map_lm, &bars[stage], k * KC, s0, b);
  };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_am) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lm) : "memory");
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_fence_init();
    issue_tma(0, 0);
    if (nk > 1) issue_tma(1, 1);
  }
  const float *lmb = p.lm + (size_t)b * S1 * C;
  const float *amb = p.am + (size_t)b * p.T * C;
  if (tid < TM) {
    const int t = t0 + tid;
    const float mx = (t < p.T) ? p.ammax[(size_t)b * p.T + t] : 0.f;
    s_ammax[tid] = mx;
    s_amneg[tid] = (t < p.T) ? -mx * kLog2e : -INFINITY;   // exp2(x*log2e - inf) = 0 for masked rows
  } else if (tid < TM + TN) {
    const int j = tid - TM, s = s0 + j;
    float lmmax = 0.f, pxlm = 0.f, pylm = 0.f, lmonly = 0.f, logus = 0.f;
    int sym = -1;
    if (s < S1) {
      lmmax = p.lmmax[(size_t)b * S1 + s];
      pylm = lmb[(size_t)s * C + p.term];
      if (s < p.S) {
        sym = p.symbols[(size_t)b * p.S + s];
        pxlm = lmb[(size_t)s * C + sym];
      }
      if (p.smoothed) {
        lmonly = logf(p.lmsum[(size_t)b * S1 + s]) + lmmax;
        logus = (sym >= 0) ? p.logu[sym] : 0.f;
      }
    }
    s_lmmax[j] = lmmax; s_lmneg[j] = (s < S1) ? -lmmax * kLog2e : -INFINITY;
    s_pxlm[j] = pxlm; s_pylm[j] = pylm; s_lmonly[j] = lmonly; s_logusym[j] = logus; s_sym[j] = sym;
  }
  if (w == 0) tmem_alloc(s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *s_tmem;
  TCT(1);

  // epilogue mapping, also used inside the k loop: thread <-> frame (TMEM lane), the four
  // warps of a lane quarter take 28 symbol columns each
  const int q = w & 3, half = w >> 2;            // `half` = column part 0..3
  const int erow = q * 32 + lane, et = t0 + erow;
  const bool t_ok = et < p.T;
  constexpr int kColsPerHalf = TN / 4;          // 28 columns per thread, 7 batches of 4
  const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
  const float py_am = __ldg(amb + (size_t)(t_ok ? et : 0) * C + p.term);
  const float amonly = (p.smoothed && t_ok) ? p.amonly[(size_t)b * p.T + et] : 0.f;
  const float logu_term = p.smoothed ? p.logu[p.term] : 0.f;
  // the symbols of this warp's columns, one per lane
  const int my_sym0 = (lane < kColsPerHalf) ? s_sym[half * kColsPerHalf + lane] : -1;
  float accr[kColsPerHalf];                     // float32 sum of the per-slice tensor-core partial sums
#pragma unroll
  for (int i = 0; i < kColsPerHalf; ++i) accr[i] = 0.f;
  auto drain_accumulator = [&]() {              // TMEM partial sums of one slice -> registers
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = half * kColsPerHalf + bi * 4;
      if (c0 < n_rows) {                        // warp-uniform
        float part[4];
        tmem_ld4(lane_addr + (uint32_t)c0, part);
#pragma unroll
        for (int e = 0; e < 4; ++e) accr[bi * 4 + e] += part[e];
      }
    }
  };
  TCT(2);
  const uint32_t idesc = umma_idesc(TM, n_rows);
  const uint32_t a_base = smem_u32(smem + kOffA), b_base = smem_u32(smem + kOffB);

  for (int k = 0; k < nk; ++k) {
    const int stage = k & 1;
    mbar_wait_bounded(&bars[stage], (uint32_t)((k >> 1) & 1));             // raw tiles landed
    if (k < 8) TCT(4 + 4 * k);
    if (k > 0) {
      mbar_wait_bounded(&bars[2], (uint32_t)((k - 1) & 1));                // previous slice's MMAs done
      tc_fence_after();
      drain_accumulator();                                                 // (operands are free again, too)
    }
    if (k < 8) TCT(5 + 4 * k);
    const float *raw_am = reinterpret_cast<const float *>(smem + kOffRaw + stage * kRawStage);
    const float *raw_lm = raw_am + TM * KC;
    const int k0 = k * KC;
    // ---- px_am[t][s] = am[b,t,sym_s]: picked out of the raw tile while it is in shared
    //      memory and parked in spare tensor-memory columns (no global gather) ----
    {
      // lane l watches column l of its warp's part; a ballot gives the hits of this slice
      uint32_t hit0 = __ballot_sync(0xffffffffu, my_sym0 >= k0 && my_sym0 < k0 + KC);
      while (hit0) {                                                       // warp-uniform loop
        const int src_lane = __ffs(hit0) - 1;
        hit0 &= hit0 - 1;
        const int sym = __shfl_sync(0xffffffffu, my_sym0, src_lane);
        const int j = half * kColsPerHalf + src_lane;
        const uint32_t v = __float_as_uint(raw_am[erow * KC + (sym - k0)]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_addr + (uint32_t)(kPxCol + j)), "r"(v)
                     : "memory");
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // ---- convert: 16-byte operand chunks (8 probabilities); three-term bf16 split by
    //      mantissa truncation: h = top 16 bits of p, m = top 16 bits of the (exact)
    //      remainder, l likewise -> h+m+l = p to 2^-24, plain ALU ops ----
    const int lim = C - k0;                     // columns of this slice that exist
    auto convert = [&](const float *raw, const float *negmax, unsigned char *ops, uint32_t stride, int nchunks) {
#pragma unroll 2
      for (int li = tid; li < nchunks; li += kThreads) {
        const int row = li >> 3, j = li & 7;
        const float *src = raw + row * KC + j * 8;
        const float nmx = negmax[row];
        const float4 v0 = *reinterpret_cast<const float4 *>(src), v1 = *reinterpret_cast<const float4 *>(src + 4);
        const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t hb[8], mb[8], lb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float pr = ex2_approx(fmaf(x[e], kLog2e, nmx));
          pr = (j * 8 + e < lim) ? pr : 0.f;
          hb[e] = __float_as_uint(pr) & 0xFFFF0000u;
          const float r1 = pr - __uint_as_float(hb[e]);
          mb[e] = __float_as_uint(r1);
          const float r2 = r1 - __uint_as_float(mb[e] & 0xFFFF0000u);
          lb[e] = __float_as_uint(r2);
        }
        unsigned char *dst = ops + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        auto pack = [](const uint32_t (&v)[8]) {   // upper halves: element e low, e+1 high
          return make_uint4(__byte_perm(v[0], v[1], 0x7632), __byte_perm(v[2], v[3], 0x7632),
                            __byte_perm(v[4], v[5], 0x7632), __byte_perm(v[6], v[7], 0x7632));
        };
        *reinterpret_cast<uint4 *>(dst) = pack(hb);
        *reinterpret_cast<uint4 *>(dst + stride) = pack(mb);
        *reinterpret_cast<uint4 *>(dst + 2 * stride) = pack(lb);
      }
    };
    convert(raw_am, s_amneg, smem + kOffA, kOpABytes, TM * 8);
    convert(raw_lm, s_lmneg, smem + kOffB, kOpBBytes, TN * 8);
    fence_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    if (k < 8) TCT(6 + 4 * k);
    if (tid == 0) {
      if (k + 2 < nk) issue_tma(k + 2, stage);   // this raw stage has been consumed
      tc_fence_after();
      // Tensor-core FP32 accumulation truncates, so (i) every slice starts from a zeroed
      // accumulator and is summed in registers, (ii) the five small products go first and
      // h*h last: <= 4 truncations at full magnitude per slice.
      const int ia[6] = {2, 0, 1, 1, 0, 0}, ib[6] = {0, 2, 1, 0, 1, 0};
      uint32_t first = 1;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const uint64_t ad = umma_desc(a_base + ia[c] * kOpABytes + ks * 32);
          const uint64_t bd = umma_desc(b_base + ib[c] * kOpBBytes + ks * 32);
          umma_bf16(tmem_d, ad, bd, idesc, first ? 0u : 1u);
          first = 0;
        }
      }
      umma_commit(&bars[2]);
    }
    if (k < 8) TCT(7 + 4 * k);
  }
  mbar_wait_bounded(&bars[2], (uint32_t)((nk - 1) & 1));
  tc_fence_after();
  drain_accumulator();
  TCT(40);

  // ---- epilogue: one frame per thread (TMEM lane), 56 symbol columns per warp half.
  //      Straight-line: everything is computed for all 8 columns of a batch, only the
  //      stores are predicated. ----
  {
    const int t = et;
    const int t_end = p.boundary[4 * b + 3];
    const float ammax = s_ammax[erow];
    float *pxb = p.px + (size_t)b * p.S * p.T1 + t;
    float *pyb = p.py + (size_t)b * S1 * p.T + t;
    const bool regular = (p.T1 == p.T + 1);
    const bool px_col_ok = t_ok || (regular && t == p.T);       // regular: column T exists and is -inf
    const bool px_inf = !t_ok || (p.rnnt_type == FRN_REGULAR && t == t_end);
    const bool smoothed = p.smoothed != 0;
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = half * kColsPerHalf + bi * 4;
      if (c0 < n_rows) {                 // warp-uniform
        float pxam[4];
        tmem_ld4(lane_addr + (uint32_t)(kPxCol + c0), pxam);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = c0 + e, s = s0 + j;
          const float norm = log_plus_tiny(accr[bi * 4 + e]) + s_lmmax[j] + ammax;
          const float py_lm = s_pylm[j], px_lm = s_pxlm[j], px_am = pxam[e];
          float py = py_am + py_lm - norm;
          float px = px_am + px_lm - norm;
          if (smoothed) {                // warp-uniform
            const float lmonly = s_lmonly[j];
            py = py * p.comb + (py_lm - lmonly) * p.lm_scale + (py_am + logu_term - amonly) * p.am_scale;
            px = px * p.comb + (px_lm - lmonly) * p.lm_scale + (px_am + s_logusym[j] - amonly) * p.am_scale;
          }
          px = px_inf ? -INFINITY : px;
          if (t_ok && s < S1) pyb[(size_t)s * p.T] = py;
          if (px_col_ok && s < p.S) pxb[(size_t)s * p.T1] = px;
        }
      }
    }
  }
  TCT(41);
  tc_fence_before();
  __syncthreads();
  TCT(42);
#ifdef FRN_TC_TIMING
  if (tid == 0 && blockIdx.x == 1 && blockIdx.z == 3) {
    printf("TC timing (cycles from start): setup %lld prologue %lld\n", tct[1] - tct[0], tct[2] - tct[0]);
    for (int k = 0; k < 8 && k < nk; ++k)
      printf("  k%d raw_wait->%lld mma_wait+drain->%lld stage+convert+sync->%lld issue->%lld\n", k, tct[4 + 4 * k] - tct[0],
             tct[5 + 4 * k] - tct[0], tct[6 + 4 * k] - tct[0], tct[7 + 4 * k] - tct[0]);
    printf("  last mma done %lld epilogue end %lld final sync %lld\n", tct[40] - tct[0], tct[41] - tct[0], tct[42] - tct[0]);
  }
#endif
  if (w == 0) tmem_dealloc(tmem_d, kTmemCols);
}

// ---------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

static bool make_map_3d(CUtensorMap *map, const float *base, int rows, int C, int B, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
  cuuint32_t box[3] = {(cuuint32_t)tc::KC, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns FRN_EUNSUPPORTED when the tensor-core path does not apply (caller falls
// back to the SIMT kernel): C % 4 != 0 (TMA needs 16-byte global strides) or
// misaligned bases.
int launch_simple_logprobs_tc(const SimpleParams &sp, cudaStream_t stream) {
  if (sp.C % 4 != 0) return FRN_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(sp.am) | reinterpret_cast<uintptr_t>(sp.lm)) & 15u) return FRN_EUNSUPPORTED;
  CUtensorMap map_am, map_lm;
  if (!make_map_3d(&map_am, sp.am, sp.T, sp.C, sp.B, tc::TM)) return FRN_EUNSUPPORTED;
  if (!make_map_3d(&map_lm, sp.lm, sp.S + 1, sp.C, sp.B, tc::TN)) return FRN_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(simple_logprobs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::kSmemBytes);
  if (e != cudaSuccess) return note_cuda_error(e);
  dim3 grid((sp.T1 + tc::TM - 1) / tc::TM, (sp.S + 1 + tc::TN - 1) / tc::TN, sp.B);
  count_launch(), simple_logprobs_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(map_am, map_lm, sp);
  return check_launch();
}

}  // namespace frn
