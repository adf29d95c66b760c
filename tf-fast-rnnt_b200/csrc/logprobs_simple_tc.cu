// Simple / smoothed log-probs (A1, A2) on the 5th-generation tensor cores.
//
//   norm[b,s,t] = log( sum_c exp(lm[b,s,c]-lmmax[b,s]) * exp(am[b,t,c]-ammax[b,t]) + tiny ) + maxes
//
// One CTA = one (utterance, 128-frame tile, <=112-symbol tile).  Per 64-wide
// slice of the vocabulary axis:
//   1. all 512 threads load their 32-byte pieces of the am and lm tiles straight from global memory
//      into registers - issued right after the previous slice's MMAs, so the loads fly while the
//      tensor core works;
//   2. they turn them into probabilities exp(x - rowmax) and split each into three bfloat16 terms
//      h+m+l (24 mantissa bits), written in the K-major SWIZZLE_128B layout tcgen05 reads, into one
//      of TWO operand stages: the conversion of slice k+1 overlaps the MMAs of slice k;
//   3. one thread issues 6 tcgen05.mma (hh, hm, mh, mm, hl, lh - everything down to 2^-24
//      relative) per 16-wide k step into one of two 128 x N float32 accumulators in tensor memory
//      and commits to the stage's mbarrier; the partial sums of slice k-2 are drained into
//      registers before its stage is overwritten (tensor-core accumulation truncates: every slice
//      starts from zero and the slices are summed in float32 registers).
// (The first version staged raw tiles through shared memory with TMA and single-buffered operands:
// convert and MMA never overlapped and the raw tiles cost a third of the kernel's shared-memory
// traffic; am[b,t,sym_s] was picked out of the raw tile - it now comes from the row-statistics
// kernel, which has every am row in flight anyway: SimpleParams::pxam_t.)
// Epilogue: tcgen05.ld the accumulator (one lattice frame per thread), then either
//   (a) log, un-shift, symbol / blank gather, smoothing terms, boundary fix-ups, and store
//       px/py coalesced along t in the reference layout (rnnt_loss.py:186-221,1290-1365)
//       - the public get_rnnt_logprobs{,_smoothed}; or
//   (b) frn_simple_loss: every live arc goes straight into the diagonal-major plane the
//       wavefront recursion streams (mi_dp.cu), as (mantissa in [1,2], integer exponent) of
//       numerator / Z with boundary masks and the delay penalty applied.  The exponent is
//       assembled from exact integers (floor parts of the shifted scores, the exponent field
//       of Z) and the mantissa from one ex2 of a fraction in [0,1): no float32 log-prob of
//       magnitude ~10 is ever formed, so an arc carries ~4e-7 instead of ~2e-6 relative
//       error and the px/py round trip + the skew kernel are gone.  The tile is transposed
//       through the (by then free) operand shared memory so that a warp writes 512
//       contiguous bytes of one diagonal; the dead remainder of the plane is filled by the
//       same launch (every CTA takes a share of the diagonals while its first TMA loads
//       are in flight).
//
// float32-accurate by construction (the occupation counts downstream need ~2^-21
// on the normaliser, which rules out plain bf16/tf32: DESIGN.md "numerics").
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "simple_params.cuh"

namespace frn {

namespace tc {
constexpr int TM = 128;   // frames per CTA  (MMA M)
constexpr int TN = 112;   // symbols per CTA (MMA N, multiple of 16)
constexpr int KC = 64;    // vocabulary slice per stage = one 128-byte swizzle row of bf16
constexpr int kThreads = 512;   // 16 warps: 4 per TMEM lane quarter, 28 symbol columns each
constexpr int kTmemCols = 256;  // two accumulators: [0,112) and [128,240)
constexpr int kAccStride = 128;
constexpr uint32_t kOpABytes = TM * KC * 2, kOpBBytes = TN * KC * 2;
// shared memory map (byte offsets from a 1024-aligned base): two operand stages, then the small block
constexpr uint32_t kOffB = 3 * kOpABytes;                      // inside a stage: 3 x A (h, m, l), then 3 x B
constexpr uint32_t kStageBytes = 3 * kOpABytes + 3 * kOpBBytes;
constexpr uint32_t kOffSmall = 2 * kStageBytes;
constexpr uint32_t kSmallBytes = 8192;
constexpr uint32_t kSmemBytes = kOffSmall + kSmallBytes + 1024;  // + alignment slack
// Accumulator columns of a thread: the four warps of a TMEM lane quarter (`half` = 0..3) take 28
// symbol columns each, 16 out of [0,64) and 12 out of [64,112): the arc-plane epilogue finishes the
// tile in two passes over those column ranges (one pass of staging fits the free shared memory).
constexpr int kColsPerHalf = TN / 4;          // 28 columns per thread, 7 batches of 4
constexpr int kPassA = 64;
__host__ __device__ constexpr int col_of(int half, int i) {
  return i < 16 ? half * 16 + i : kPassA + half * 12 + (i - 16);
}
constexpr double kLog2eD = 1.4426950408889634074;
constexpr float kLog2eLo = 1.92596299112661746e-8f;   // log2(e) - (float)log2(e)
struct Small {                                        // per-CTA row / column constants
  double sm_x[TN], sm_y[TN];                          // smoothed: lm_scale * log2e * (lm[s,sym|blank] - lmonly[s])
  uint64_t bars[4];                                   // full[2], done[2] (one each per operand stage)
  float amneg[TM], lmneg[TN];                         // -max * log2e (-inf: row masked)
  float ammax[TM], lmmax[TN];
  float pxlm[TN], pylm[TN], lmonly[TN], logusym[TN];
  int sym[TN];
  // arc-plane epilogue: (lm[s,sym] - lmmax) * log2e and (lm[s,blank] - lmmax) * log2e as integer + fraction in [0,1)
  int bx_e[TN], by_e[TN];
  float bx_f[TN], by_f[TN];
  uint32_t tmem;
};
static_assert(sizeof(Small) <= kSmallBytes, "small shared-memory block");

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

template <bool kXY>
__global__ void __launch_bounds__(tc::kThreads, 1)
simple_logprobs_tc_kernel(SimpleParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B operands need a 1024-byte aligned base; keep the arithmetic on the
  // __shared__ array so that accesses stay LDS/STS (a pointer rebuilt from an integer
  // degrades them to generic LD/ST)
  unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.z, t0 = blockIdx.x * TM, s0 = blockIdx.y * TN;
  const int S1 = p.S + 1, C = p.C;
  const int nk = (C + KC - 1) / KC;
  const int n_rows = min(TN, round_up(S1 - s0, 16));   // MMA N for this tile (multiple of 16)
  Small &sm = *reinterpret_cast<Small *>(smem + kOffSmall);
  uint64_t *bars = sm.bars;

  // utterance geometry (arc-plane output only)
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const bool bd_ok = bd.z - bd.x >= 0 && bd.w - bd.y >= 0 && bd.x >= 0 && bd.y >= 0 && bd.z <= p.S && bd.w <= p.T;
  const int s_begin = bd.x, t_begin = bd.y, Sb = bd_ok ? bd.z - bd.x : -1, Tb = bd_ok ? bd.w - bd.y : -1;
  const float2 dead2 = make_float2(0.f, __int_as_float(kNegI));
  float4 *XYb = kXY ? p.XY + (size_t)b * p.Dn * p.P : nullptr;
  // a tile that holds no live arc has nothing to contract
  const bool tile_dead = kXY && (!bd_ok || t0 >= bd.w || t0 + TM <= t_begin || s0 > bd.z || s0 + TN <= s_begin);

  if (tid == 0 && !tile_dead) {
    mbar_init(&bars[0], kThreads / 32 - 1); mbar_init(&bars[1], kThreads / 32 - 1);    // full[2]: one arrival per converter warp
    mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);                                    // done[2]: the tcgen05.commit
    mbar_fence_init();
  }
  if constexpr (kXY) {
    // Dead remainder of the arc plane.  Half X of cell (d, r) is the symbol arc from lattice row r-1 at
    // relative frame tx, half Y the blank arc of row r at relative frame ty; an arc is live iff its source
    // lies inside the boundary box (mutual_information_cuda.cu:295-303 load rules; rnnt_loss.py:51-60 makes
    // frame t_end dead).  Live halves are written by the epilogue below, all others here: the tiles of an
    // utterance share its diagonals evenly.  Only diagonals the recursion streams are touched.
    const int ntile = gridDim.x * gridDim.y, lt = blockIdx.y * gridDim.x + blockIdx.x;
    const int Dfill = bd_ok ? min(p.Dn, round_up(Tb + p.k * Sb + 1, kChunk)) : 0;
    const int dlo = (int)((long long)Dfill * lt / ntile), dhi = (int)((long long)Dfill * (lt + 1) / ntile);
    const float4 dead4 = make_float4(dead2.x, dead2.y, dead2.x, dead2.y);
    for (int d = dlo + w; d < dhi; d += kThreads / 32)
      for (int r = lane; r < p.P; r += 32) {
        const int ty = d - 1 - p.k * r, tx = ty + p.k;
        const bool xa = r >= 1 && r <= Sb && tx >= 0 && tx < Tb;
        const bool ya = r <= Sb && ty >= 0 && ty < Tb;
        float4 *dst = XYb + (size_t)d * p.P + r;
        if (!xa && !ya) *dst = dead4;
        else if (!xa) *reinterpret_cast<float2 *>(dst) = dead2;
        else if (!ya) *(reinterpret_cast<float2 *>(dst) + 1) = dead2;
      }
    if (tile_dead) return;
  }
  const float *lmb = p.lm + (size_t)b * S1 * C;
  const float *amb = p.am + (size_t)b * p.T * C;
  if (tid < TM) {
    const int t = t0 + tid;
    const float mx = (t < p.T) ? p.ammax[(size_t)b * p.T + t] : 0.f;
    sm.ammax[tid] = mx;
    sm.amneg[tid] = (t < p.T) ? -mx * kLog2e : -INFINITY;   // exp2(x*log2e - inf) = 0 for masked rows
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
    sm.lmmax[j] = lmmax; sm.lmneg[j] = (s < S1) ? -lmmax * kLog2e : -INFINITY;
    sm.pxlm[j] = pxlm; sm.pylm[j] = pylm; sm.lmonly[j] = lmonly; sm.logusym[j] = logus; sm.sym[j] = sym;
    if constexpr (kXY) {
      // shifted lm scores in log2 units, integer + fraction (formed once per column in float64)
      auto split = [](float x, float mx, int &e, float &f) {
        const double v = ((double)x - (double)mx) * kLog2eD;
        if (!(v > -1.0e8)) { e = kNegI; f = 0.f; return; }
        const double fl = floor(v);
        e = (int)fl; f = (float)(v - fl);
      };
      split(pxlm, lmmax, sm.bx_e[j], sm.bx_f[j]);
      split(pylm, lmmax, sm.by_e[j], sm.by_f[j]);
      sm.sm_x[j] = (double)p.lm_scale * kLog2eD * ((double)pxlm - (double)lmonly);
      sm.sm_y[j] = (double)p.lm_scale * kLog2eD * ((double)pylm - (double)lmonly);
    }
  }
  if (w == 0) tmem_alloc(&sm.tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sm.tmem;

  // epilogue mapping, also used inside the k loop: thread <-> frame (TMEM lane), the four
  // warps of a lane quarter take 28 symbol columns each (col_of)
  const int q = w & 3, half = w >> 2;            // `half` = column part 0..3
  const int erow = q * 32 + lane, et = t0 + erow;
  const bool t_ok = et < p.T;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
  const float py_am = __ldg(amb + (size_t)(t_ok ? et : 0) * C + p.term);
  const float amonly = (p.smoothed && t_ok) ? p.amonly[(size_t)b * p.T + et] : 0.f;
  const float logu_term = p.smoothed ? p.logu[p.term] : 0.f;
  // am[b,t,sym_s] for this thread's frame: gathered by the row-statistics kernel while it had the row in flight
  const float *pxam_row = p.pxam_t + ((size_t)b * p.T + (t_ok ? et : 0)) * p.S;
  float accr[kColsPerHalf];                     // float32 sum of the per-slice tensor-core partial sums
#pragma unroll
  for (int i = 0; i < kColsPerHalf; ++i) accr[i] = 0.f;
  auto drain_accumulator = [&](int acc) {       // TMEM partial sums of one slice -> registers
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = col_of(half, bi * 4);
      if (c0 < n_rows) {                        // warp-uniform
        float part[4];
        tmem_ld4(lane_addr + (uint32_t)(acc * kAccStride + c0), part);
#pragma unroll
        for (int e = 0; e < 4; ++e) accr[bi * 4 + e] += part[e];
      }
    }
  };
  const uint32_t idesc = umma_idesc(TM, n_rows);

  // ---- roles.  Warps 0..14 (480 threads) convert: 1024 pieces of the am tile + 896 of the lm tile = 1920 = 4 per
  //      thread (a piece = 8 consecutive values of one row: 32 bytes in, 3 x 16 bytes out).  Warp 15 issues the
  //      MMAs: a thread that also converted made every slice wait for its 24 tcgen05.mma behind its share of the
  //      conversion (23 % of all warp time sat in the per-slice block barrier).  There is no block barrier in
  //      the loop: converters -> issuer through full[stage] (one arrival per converter warp), issuer -> everybody
  //      through the tcgen05.commit on done[stage]. ----
  constexpr int kConvWarps = kThreads / 32 - 1, kConvThreads = kConvWarps * 32, kPieces = (TM + TN) * 8 / kConvThreads;
  static_assert(kPieces * kConvThreads == (TM + TN) * 8, "pieces divide evenly over the converter threads");
  const bool issuer = (w == kConvWarps);
  uint64_t *full = bars, *done = bars + 2;
  float4 rr[kPieces][2];
  // piece i of this thread -> (tile, row, 16-byte chunk); warp-uniform in `is_a`
  auto piece = [&](int i, bool &is_a, int &row, int &j) {
    int pi = tid + i * kConvThreads;
    is_a = pi < TM * 8;
    if (!is_a) pi -= TM * 8;
    row = pi >> 3; j = pi & 7;
  };
  // per piece, once: where its row starts (element offset inside the utterance's am / lm, -1: row outside the
  // tensor), where its operand chunk goes, and its row's -max * log2e
  int src_off[kPieces];
  uint32_t dst_off[kPieces];
  float nmx_of[kPieces];
#pragma unroll
  for (int i = 0; i < kPieces; ++i) {
    bool is_a; int row, j;
    piece(i, is_a, row, j);
    const bool row_ok = is_a ? (t0 + row < p.T) : (s0 + row < S1);
    src_off[i] = row_ok ? (is_a ? t0 + row : s0 + row) * C + j * 8 : -1;
    dst_off[i] = (is_a ? 0u : kOffB) + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((j ^ (row & 7)) << 4);
    nmx_of[i] = issuer ? 0.f : (is_a ? sm.amneg[row] : sm.lmneg[row]);      // masked rows: -inf -> exp2 = 0
  }
  auto load_slice = [&](int k) {
    const int k0 = k * KC;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const bool is_a = tid + i * kConvThreads < TM * 8;
      const int c = k0 + ((tid + i * kConvThreads) & 7) * 8;
      rr[i][0] = rr[i][1] = z;                    // rows / columns outside the tensors stay zero
      if (src_off[i] >= 0) {
        const float4 *src = reinterpret_cast<const float4 *>((is_a ? amb : lmb) + src_off[i] + k0);
        if (c < C) rr[i][0] = __ldg(src);         // C % 4 == 0: a float4 is inside or outside the row
        if (c + 4 < C) rr[i][1] = __ldg(src + 1);
      }
    }
  };
  // ---- registers -> operand stage: exp, three-term bf16 split by mantissa truncation (h = top 16 bits of
  //      p, m = top 16 bits of the exact remainder, l likewise: h+m+l = p to 2^-24, plain ALU ops),
  //      16-byte chunks in the K-major SWIZZLE_128B layout ----
  auto store_piece = [&](auto masked, const float4 (&r)[2], float nmx, int lim, unsigned char *dst, uint32_t stride) {
    const float x[8] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w};
    uint32_t hb[8], mb[8], lb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float pr = ex2_approx(fmaf(x[e], kLog2e, nmx));
      if (decltype(masked)::value) pr = (e < lim) ? pr : 0.f;    // only the last slice has columns beyond C
      hb[e] = __float_as_uint(pr) & 0xFFFF0000u;
      const float r1 = pr - __uint_as_float(hb[e]);
      mb[e] = __float_as_uint(r1);
      const float r2 = r1 - __uint_as_float(mb[e] & 0xFFFF0000u);
      lb[e] = __float_as_uint(r2);
    }
    auto pack = [](const uint32_t (&v)[8]) {   // upper halves: element e low, e+1 high
      return make_uint4(__byte_perm(v[0], v[1], 0x7632), __byte_perm(v[2], v[3], 0x7632),
                        __byte_perm(v[4], v[5], 0x7632), __byte_perm(v[6], v[7], 0x7632));
    };
    *reinterpret_cast<uint4 *>(dst) = pack(hb);
    *reinterpret_cast<uint4 *>(dst + stride) = pack(mb);
    *reinterpret_cast<uint4 *>(dst + 2 * stride) = pack(lb);
  };
  auto store_slice = [&](int k, unsigned char *stage) {
    const int k0 = k * KC;
    const bool full_slice = k0 + KC <= C;        // block-uniform: no column of this slice lies beyond C
    if (full_slice) {
#pragma unroll
      for (int i = 0; i < kPieces; ++i)
        store_piece(std::false_type{}, rr[i], nmx_of[i], 8, stage + dst_off[i],
                    tid + i * kConvThreads < TM * 8 ? kOpABytes : kOpBBytes);
    } else {
#pragma unroll
      for (int i = 0; i < kPieces; ++i)
        store_piece(std::true_type{}, rr[i], nmx_of[i], C - k0 - ((tid + i * kConvThreads) & 7) * 8, stage + dst_off[i],
                    tid + i * kConvThreads < TM * 8 ? kOpABytes : kOpBBytes);
    }
  };

  if (!issuer) load_slice(0);
  for (int k = 0; k < nk; ++k) {
    const int st = k & 1;
    unsigned char *stage = smem + (uint32_t)st * kStageBytes;
    if (k >= 2) {                                // slice k-2 used this stage and this accumulator
      mbar_wait_bounded(&done[st], (uint32_t)(((k >> 1) - 1) & 1));
      tc_fence_after();
      drain_accumulator(st);
      tc_fence_before();                         // the drain's tcgen05.ld before the arrival below / the next MMAs
    }
    if (!issuer) {
      store_slice(k, stage);
      if (k + 1 < nk) load_slice(k + 1);         // registers are free again: in flight until the next conversion
      fence_async_smem();                        // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    } else {
      mbar_wait_bounded(&full[st], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      if (lane == 0) {
        // Tensor-core FP32 accumulation truncates, so (i) every slice starts from a zeroed
        // accumulator and is summed in registers, (ii) the five small products go first and
        // h*h last: <= 4 truncations at full magnitude per slice.
        const uint32_t a_base = smem_u32(stage), b_base = smem_u32(stage + kOffB);
        const uint32_t acc = tmem_d + (uint32_t)(st * kAccStride);
        const int ia[6] = {2, 0, 1, 1, 0, 0}, ib[6] = {0, 2, 1, 0, 1, 0};
        uint32_t first = 1;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t ad = umma_desc(a_base + ia[c] * kOpABytes + ks * 32);
            const uint64_t bd = umma_desc(b_base + ib[c] * kOpBBytes + ks * 32);
            umma_bf16(acc, ad, bd, idesc, first ? 0u : 1u);
            first = 0;
          }
        }
        umma_commit(&done[st]);
      }
      __syncwarp();
    }
  }
  // the last two slices (in order: the sums are formed in slice order whatever the timing)
  for (int k = max(nk - 2, 0); k < nk; ++k) {
    mbar_wait_bounded(&done[k & 1], (uint32_t)((k >> 1) & 1));
    tc_fence_after();
    drain_accumulator(k & 1);
  }
  tc_fence_before();
  __syncthreads();                               // every warp has left the operand stages: the epilogue reuses them
  tc_fence_after();

  if constexpr (!kXY) {
    // ---- epilogue (a): one frame per thread (TMEM lane), 28 symbol columns per thread.
    //      Straight-line: everything is computed for all 4 columns of a batch, only the
    //      stores are predicated. ----
    const int t = et;
    const int t_end = bd.w;
    const float ammax = sm.ammax[erow];
    float *pxb = p.px + (size_t)b * p.S * p.T1 + t;
    float *pyb = p.py + (size_t)b * S1 * p.T + t;
    const bool regular = (p.T1 == p.T + 1);
    const bool px_col_ok = t_ok || (regular && t == p.T);       // regular: column T exists and is -inf
    const bool px_inf = !t_ok || (p.rnnt_type == FRN_REGULAR && t == t_end);
    const bool smoothed = p.smoothed != 0;
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = col_of(half, bi * 4);
      if (c0 < n_rows) {                 // warp-uniform
        float pxam[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pxam[e] = (t_ok && s0 + c0 + e < p.S) ? pxam_row[s0 + c0 + e] : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = c0 + e, s = s0 + j;
          const float norm = log_plus_tiny(accr[bi * 4 + e]) + sm.lmmax[j] + ammax;
          const float py_lm = sm.pylm[j], px_lm = sm.pxlm[j], px_am = pxam[e];
          float py = py_am + py_lm - norm;
          float px = px_am + px_lm - norm;
          if (smoothed) {                // warp-uniform
            const float lmonly = sm.lmonly[j];
            py = py * p.comb + (py_lm - lmonly) * p.lm_scale + (py_am + logu_term - amonly) * p.am_scale;
            px = px * p.comb + (px_lm - lmonly) * p.lm_scale + (px_am + sm.logusym[j] - amonly) * p.am_scale;
          }
          px = px_inf ? -INFINITY : px;
          if (t_ok && s < S1) pyb[(size_t)s * p.T] = py;
          if (px_col_ok && s < p.S) pxb[(size_t)s * p.T1] = px;
        }
      }
    }
  } else {
    // ---- epilogue (b): arcs into the diagonal-major plane.  Thread (frame e = erow, column j) holds
    //      symbol arc X = px[s0+j][t0+e] and blank arc Y = py[s0+j][t0+e]; both belong to plane diagonal
    //      d = Delta + 1 + e + k*j (Delta = t0 - t_begin + k*(s0 - s_begin)), X to row s'+1, Y to row s'.
    //      Two passes (columns [0,64), [64,112)): stage [column][frame] in the free operand memory,
    //      then every warp writes whole diagonals, lanes along the rows. ----
    const int kk = p.k;
    const int pitch = 129 - kk;                 // (pitch - k) odd: the skewed read-out is bank-conflict free
    float2 *stX = reinterpret_cast<float2 *>(smem);
    float2 *stY = stX + kPassA * 129;
    const float ammax = sm.ammax[erow];
    const bool smoothed = p.smoothed != 0;
    // per-frame constants: blank score and delay penalty in log2 units, integer + fraction
    int ay_e = kNegI, pen_e = 0;
    float ay_f = 0.f, pen_f = 0.f;
    double pen_d = 0.0, amy_d = 0.0;
    {
      const double v = ((double)py_am - (double)ammax) * kLog2eD;
      if (v > -1.0e8) { const double fl = floor(v); ay_e = (int)fl; ay_f = (float)(v - fl); }
      if (p.delay_penalty != 0.f) {             // rnnt_loss.py:316-321: float64, rounded to float32, added to px
        pen_d = (double)delay_penalty_value(bd.w, et, p.delay_penalty) * kLog2eD;
        const double fl = floor(pen_d);
        pen_e = (int)fl; pen_f = (float)(pen_d - fl);
      }
      if (smoothed) amy_d = (double)p.am_scale * kLog2eD * ((double)py_am + (double)logu_term - (double)amonly);
    }
    const double comb_d = (double)p.comb, amsc_d = (double)p.am_scale * kLog2eD;
    auto finish = [&](int e_int, float f) {     // exponent + fraction in (-2, 4) -> (mantissa in [1,2], exponent)
      const float kf = floorf(f);
      const int ex = e_int + (int)kf;
      return (ex > kNegIThresh) ? make_float2(ex2_approx(f - kf), __int_as_float(ex)) : dead2;
    };
    auto finish_d = [&](double v) {
      if (!(v > -1.0e8)) return dead2;
      const double fl = floor(v);
      return make_float2(ex2_approx((float)(v - fl)), __int_as_float((int)fl));
    };
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int jb = pass ? kPassA : 0, ncols = pass ? TN - kPassA : kPassA;
      if (jb < n_rows) {                        // block-uniform
#pragma unroll
        for (int bi = (pass ? 4 : 0); bi < (pass ? 7 : 4); ++bi) {
          const int c0 = col_of(half, bi * 4);
          if (c0 < n_rows) {                    // warp-uniform
            float pxam[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pxam[e] = (t_ok && s0 + c0 + e < p.S) ? pxam_row[s0 + c0 + e] : 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c0 + e;
              // log2 Z = exponent field + lg2(mantissa); Z < FLT_MIN counts as the reference's log(0 + tiny)
              const float z = accr[bi * 4 + e];
              const uint32_t u = __float_as_uint(z);
              const bool ztiny = z < 1.1754944e-38f;
              const int ze = ztiny ? -149 : (int)(u >> 23) - 127;
              const float lg = ztiny ? 0.f : lg2_approx(__uint_as_float((u & 0x007FFFFFu) | 0x3F800000u));
              // (am[t,sym] - ammax) * log2e: product split exactly into head + tail
              const float a = pxam[e] - ammax;
              const float hi = a * kLog2e;
              const float lo = fmaf(a, kLog2e, -hi) + a * kLog2eLo;
              const float hfl = floorf(hi);
              const bool x_ok = a > -1.0e8f;
              const int ex_i = (x_ok ? (int)hfl : kNegI) + sm.bx_e[j] - ze;
              const float fx = ((hi - hfl) + lo) + sm.bx_f[j] - lg;
              const int ey_i = ay_e + sm.by_e[j] - ze;
              const float fy = ay_f + sm.by_f[j] - lg;
              float2 X, Y;
              if (!smoothed) {                  // warp-uniform
                X = finish(ex_i + pen_e, x_ok ? fx + pen_f : 0.f);
                Y = finish(ey_i, fy);
              } else {                          // rnnt_loss.py:1342-1360 in log2 units
                const double vx = comb_d * ((double)ex_i + (double)fx) + sm.sm_x[j] +
                                  amsc_d * ((double)pxam[e] + (double)sm.logusym[j] - (double)amonly) + pen_d;
                const double vy = comb_d * ((double)ey_i + (double)fy) + sm.sm_y[j] + amy_d;
                X = (x_ok && ex_i > kNegIThresh) ? finish_d(vx) : dead2;
                Y = (ey_i > kNegIThresh) ? finish_d(vy) : dead2;
              }
              stX[(j - jb) * pitch + erow] = X;
              stY[(j - jb) * pitch + erow] = Y;
            }
          }
        }
      }
      __syncthreads();
      if (jb < n_rows) {
        const int rho = s0 + jb - s_begin, Delta = (t0 - t_begin) + kk * rho;
        const int ND = TM + kk * (ncols - 1), ngrp = (ncols + 1 + 31) / 32;
        for (int delta = w; delta < ND; delta += kThreads / 32) {
          const int d = Delta + 1 + delta;
          for (int g = 0; g < ngrp; ++g) {
            const int jp = g * 32 + lane, r = rho + jp;
            // same liveness rule as the fill above, in plane coordinates
            const int ty = d - 1 - kk * r, tx = ty + kk;
            const int ex_ = delta - kk * (jp - 1), ey_ = delta - kk * jp;     // staged frame of either half
            const bool xa = jp >= 1 && jp <= ncols && ex_ >= 0 && ex_ < TM && r >= 1 && r <= Sb && tx >= 0 && tx < Tb;
            const bool ya = jp < ncols && ey_ >= 0 && ey_ < TM && r >= 0 && r <= Sb && ty >= 0 && ty < Tb;
            float2 X = dead2, Y = dead2;
            if (xa) X = stX[(jp - 1) * pitch + ex_];
            if (ya) Y = stY[jp * pitch + ey_];
            float4 *dst = XYb + (size_t)d * p.P + r;
            if (xa && ya) *dst = make_float4(X.x, X.y, Y.x, Y.y);
            else if (xa) *reinterpret_cast<float2 *>(dst) = X;
            else if (ya) *(reinterpret_cast<float2 *>(dst) + 1) = Y;
          }
        }
      }
      if (pass == 0) __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) tmem_dealloc(tmem_d, kTmemCols);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool simple_logprobs_tc_applicable(const float *lm, const float *am, int C) {
  // 128-bit loads of 8-column pieces: C % 4 == 0 and 16-byte aligned bases
  return C % 4 == 0 && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm)) & 15u) == 0;
}

// returns FRN_EUNSUPPORTED when the tensor-core path does not apply (caller falls back to the SIMT kernel):
// C % 4 != 0 or misaligned bases.  sp.XY != nullptr selects the arc-plane output (regular / modified only).
int launch_simple_logprobs_tc(const SimpleParams &sp, cudaStream_t stream) {
  if (!simple_logprobs_tc_applicable(sp.lm, sp.am, sp.C) || sp.pxam_t == nullptr) return FRN_EUNSUPPORTED;
  if (sp.XY != nullptr && sp.rnnt_type == FRN_CONSTRAINED) return FRN_EUNSUPPORTED;
  auto kernel = sp.XY ? simple_logprobs_tc_kernel<true> : simple_logprobs_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
  if (e != cudaSuccess) return note_cuda_error(e);
  // arc-plane output: frames 0..T-1 carry arcs (the regular lattice's extra column T has none)
  dim3 grid(((sp.XY ? sp.T : sp.T1) + tc::TM - 1) / tc::TM, (sp.S + 1 + tc::TN - 1) / tc::TN, sp.B);
  count_launch(), kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(sp);
  return check_launch();
}

}  // namespace frn
