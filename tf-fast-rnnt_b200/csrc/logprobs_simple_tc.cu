// Simple / smoothed log-probs (A1, A2) on the 5th-generation tensor cores.
//
//   norm[b,s,t] = log( sum_c exp(lm[b,s,c]-lmmax[b,s]) * exp(am[b,t,c]-ammax[b,t]) + tiny ) + maxes
//
// One CTA = one (utterance, 128-frame tile, <=112-symbol tile).  The operands are NOT built here: the
// row-statistics kernel (logprobs_simple.cu), which streams every am / lm row anyway, leaves each probability
// p = exp(x - rowmax) * 2^15 as a two-term float16 split  p = h + l * 2^-11  (h = fp16(p), l = fp16((p - h) 2^11):
// 22 mantissa bits, and l has the range of h, so small probabilities keep their relative precision) - once per
// element instead of once per tile (the in-kernel conversion was MUFU- and issue-bound: every CTA re-exponentiated
// and re-split its 240 x 64 slice, 6 x redundant over the grid at the c4 shape).  What is left is a GEMM pipeline:
//   warp 16   one thread: per 64-wide vocabulary slice four TMA boxes (A_h, A_l: 128 frames; B_h, B_l: 112 symbols;
//             SWIZZLE_128B, i.e. exactly the K-major layout tcgen05 reads; rows / columns outside the tensors
//             arrive as zeros) into one of THREE 60 KB stages, full[stage] counts the bytes;
//   warp 17   one thread: per 16-wide k step three tcgen05.mma - l_a h_b and h_a l_b into the cross-term
//             accumulator (kept over all slices, weight 2^-11), h_a h_b into one of two per-slice accumulators;
//             tcgen05.commit frees the stage (empty[stage]) and publishes the slice (accfull[acc]);
//   warps 0-15 drain the per-slice accumulator into float32 registers (tensor-core accumulation truncates: a
//             slice starts from zero and the slices are summed in registers) and hand it back (accfree[acc]).
// Dropped: l_a l_b (2^-22 relative).  Z = (sum h h + 2^-11 sum cross) * 2^-30.
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
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "simple_params.cuh"

namespace frn {
#ifdef FRN_TC_TIMING
// diagnostic build only: clock of block (1,0,0), thread 0 at the phase boundaries of the normaliser
__device__ long long g_tc_timing[8];
#define FRN_TCT(i) if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) g_tc_timing[(i) + 1] = clock64();
#else
#define FRN_TCT(i)
#endif

namespace tc {
constexpr int TM = 128;   // frames per CTA  (MMA M)
constexpr int TN = 112;   // symbols per CTA (MMA N, multiple of 16)
constexpr int KC = 64;    // vocabulary slice per stage = one 128-byte swizzle row of float16
constexpr int kEpiWarps = 16;   // drain / epilogue warps: 4 per TMEM lane quarter, 28 symbol columns each
constexpr int kThreads = (kEpiWarps + 2) * 32;   // + the TMA producer warp + the MMA issuer warp
constexpr int kStages = 3;
constexpr int kTmemCols = 512;  // two per-slice accumulators [0,112), [128,240) and the cross-term accumulator [256,368)
constexpr int kAccStride = 128;
constexpr int kCrossCol = 256;
constexpr uint32_t kOpABytes = TM * KC * 2, kOpBBytes = TN * KC * 2;
// shared memory map (byte offsets from a 1024-aligned base): three operand stages, then the small block
constexpr uint32_t kOffB = 2 * kOpABytes;                      // inside a stage: A_h, A_l, then B_h, B_l
constexpr uint32_t kStageBytes = 2 * kOpABytes + 2 * kOpBBytes;
constexpr uint32_t kOffSmall = kStages * kStageBytes;
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
  uint64_t bars[2 * kStages + 4];                     // full[kStages], empty[kStages], accfull[2], accfree[2]
  float ammax[TM], lmmax[TN];
  float lmsum_g[TN];                                  // sum_c exp(lm - lmmax) of the column's row (accuracy guard)
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
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (bit 4), A=B=f16 (format 0), both K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
// four accumulator columns of this thread's lane; the caller waits once (tmem_ld_wait) for a batch of loads
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
simple_logprobs_tc_kernel(const __grid_constant__ CUtensorMap map_amh, const __grid_constant__ CUtensorMap map_aml,
                          const __grid_constant__ CUtensorMap map_lmh, const __grid_constant__ CUtensorMap map_lml,
                          SimpleParams p) {
  using namespace tc;
#ifdef FRN_TC_TIMING
  if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) g_tc_timing[0] = clock64();
#endif
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

  uint64_t *full = bars, *empty = bars + kStages, *accfull = bars + 2 * kStages, *accfree = accfull + 2;
  if (tid == 0 && !tile_dead) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&accfull[0], 1); mbar_init(&accfull[1], 1);                      // the tcgen05.commit
    mbar_init(&accfree[0], kEpiWarps); mbar_init(&accfree[1], kEpiWarps);      // one arrival per drain warp
    mbar_fence_init();
  }
  __syncthreads();
  // ---- TMA producer (warp 16, one thread): started before the dead-plane fill and the per-column constants, so
  //      the first stages are in flight while the other warps set up.  Box coordinates: (vocabulary column,
  //      frame / symbol row, utterance). ----
  auto produce = [&](int k) {
    const int st = k % kStages;
    unsigned char *stage = smem + (uint32_t)st * kStageBytes;
    if (k >= kStages) mbar_wait_bounded(&empty[st], (uint32_t)((k / kStages - 1) & 1));
    mbar_arrive_expect_tx(&full[st], kStageBytes);
    tma_load_3d(stage, &map_amh, &full[st], k * KC, t0, b);
    tma_load_3d(stage + kOpABytes, &map_aml, &full[st], k * KC, t0, b);
    tma_load_3d(stage + kOffB, &map_lmh, &full[st], k * KC, s0, b);
    tma_load_3d(stage + kOffB + kOpBBytes, &map_lml, &full[st], k * KC, s0, b);
  };
  const bool producer = (w == kEpiWarps), issuer = (w == kEpiWarps + 1);
  if (producer && lane == 0 && !tile_dead)
    for (int k = 0; k < min(nk, kStages); ++k) produce(k);
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
    for (int d = dlo + w; d < dhi; d += kThreads / 32) {
      // Rows whose two arcs are both live need nothing here, and for a fixed diagonal they form ONE interval
      // [lo, hi] (k = 1: 0 <= d-1-r and d-r < Tb with 1 <= r <= Sb; k = 0: all of 1..Sb or none): the lanes walk
      // the other rows only - a third of the plane at the c2 shape.
      const int lo = p.k ? max(d - Tb + 1, 1) : 1;
      const int hi = p.k ? min(d - 1, Sb) : ((d - 1 >= 0 && d - 1 < Tb) ? Sb : 0);
      const int len = max(hi - lo + 1, 0);
      for (int idx = lane; idx < p.P - len; idx += 32) {
        const int r = (len > 0 && idx >= lo) ? idx + len : idx;
        const int ty = d - 1 - p.k * r, tx = ty + p.k;
        const bool xa = r >= 1 && r <= Sb && tx >= 0 && tx < Tb;
        const bool ya = r <= Sb && ty >= 0 && ty < Tb;
        float4 *dst = XYb + (size_t)d * p.P + r;
        if (!xa && !ya) *dst = dead4;
        else if (!xa) *reinterpret_cast<float2 *>(dst) = dead2;
        else if (!ya) *(reinterpret_cast<float2 *>(dst) + 1) = dead2;
      }
    }
    if (tile_dead) return;
  }
  if (tid < TM) {
    const int t = t0 + tid;
    const float mx = (t < p.T) ? p.ammax[(size_t)b * p.T + t] : 0.f;
    sm.ammax[tid] = mx;
  } else if (tid < TM + TN) {
    const int j = tid - TM, s = s0 + j;
    float lmmax = 0.f, pxlm = 0.f, pylm = 0.f, lmonly = 0.f, logus = 0.f;
    int sym = -1;
    if (s < S1) {
      lmmax = p.lmmax[(size_t)b * S1 + s];
      pylm = p.gat.lm_term[(size_t)b * S1 + s];
      if (s < p.S) {
        sym = p.symbols[(size_t)b * p.S + s];
        pxlm = p.gat.lm_sym[(size_t)b * S1 + s];
      }
      if (p.smoothed) {
        lmonly = logf(p.lmsum[(size_t)b * S1 + s]) + lmmax;
        logus = (sym >= 0) ? p.logu[sym] : 0.f;
      }
    }
    sm.lmmax[j] = lmmax;
    sm.lmsum_g[j] = (s < S1) ? p.gat.lm_sum[(size_t)b * S1 + s] : 0.f;
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
  FRN_TCT(0)
  tc_fence_after();
  const uint32_t tmem_d = sm.tmem;

  // epilogue mapping, also used inside the k loop: thread <-> frame (TMEM lane), the four
  // warps of a lane quarter take 28 symbol columns each (col_of); warps 16 / 17 (TMA, MMA issue) hold no columns
  const bool epi = w < kEpiWarps;
  const int q = w & 3, half = epi ? (w >> 2) : 0;  // `half` = column part 0..3
  const int erow = q * 32 + lane, et = t0 + erow;
  const bool t_ok = et < p.T;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
  const float py_am = __ldg(p.gat.am_term + (size_t)b * p.T + (t_ok ? et : 0));
  const float sa = t_ok ? __ldg(p.gat.am_sum + (size_t)b * p.T + et) : 0.f;   // accuracy guard (after the k loop)
  const float amonly = (p.smoothed && t_ok) ? p.amonly[(size_t)b * p.T + et] : 0.f;
  const float logu_term = p.smoothed ? p.logu[p.term] : 0.f;
  // am[b,t,sym_s] for this thread's frame: gathered by the row-statistics kernel while it had the row in flight
  const float *pxam_row = p.pxam_t + ((size_t)b * p.T + (t_ok ? et : 0)) * p.S;
  float accr[kColsPerHalf];                     // float32 sum of the per-slice tensor-core partial sums
#pragma unroll
  for (int i = 0; i < kColsPerHalf; ++i) accr[i] = 0.f;
  // run-time indexed access for the guard's rare path (a select chain: accr stays in registers)
  auto accr_at = [&](int i) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kColsPerHalf; ++k) v = (k == i) ? accr[k] : v;
    return v;
  };
  auto accr_set = [&](int i, float v) {
#pragma unroll
    for (int k = 0; k < kColsPerHalf; ++k) accr[k] = (k == i) ? v : accr[k];
  };
  // TMEM columns [col, col + 112) of this thread's lane -> registers: all loads issued, one wait
  auto drain_accumulator = [&](uint32_t col, auto &&fold) {
    uint32_t part[kColsPerHalf / 4][4];
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = col_of(half, bi * 4);
      if (c0 < n_rows) tmem_ld4_nowait(lane_addr + col + (uint32_t)c0, part[bi]);      // warp-uniform
    }
    tmem_ld_wait();
#pragma unroll
    for (int bi = 0; bi < kColsPerHalf / 4; ++bi) {
      const int c0 = col_of(half, bi * 4);
      if (c0 < n_rows) {
#pragma unroll
        for (int e = 0; e < 4; ++e) fold(accr[bi * 4 + e], __uint_as_float(part[bi][e]));
      }
    }
  };
  const uint32_t idesc = umma_idesc(TM, n_rows);

  if (producer) {
    if (lane == 0)
      for (int k = kStages; k < nk; ++k) produce(k);
    __syncwarp();
  } else if (issuer) {
    if (lane == 0) {
      for (int k = 0; k < nk; ++k) {
        const int st = k % kStages, ab = k & 1;
        mbar_wait_bounded(&full[st], (uint32_t)((k / kStages) & 1));
        if (k >= 2) mbar_wait_bounded(&accfree[ab], (uint32_t)(((k >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t a_h = smem_u32(smem + (uint32_t)st * kStageBytes), a_l = a_h + kOpABytes;
        const uint32_t b_h = a_h + kOffB, b_l = b_h + kOpBBytes;
        const uint32_t acc = tmem_d + (uint32_t)(ab * kAccStride), cross = tmem_d + (uint32_t)kCrossCol;
        // Tensor-core FP32 accumulation truncates: the full-magnitude products h_a h_b start from a zeroed
        // accumulator every slice (4 truncations, then float32 adds in registers); the cross terms weigh
        // 2^-11 and stay in tensor memory over all slices.
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks)
          umma_f16(cross, umma_desc(a_l + ks * 32), umma_desc(b_h + ks * 32), idesc, (k > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks)
          umma_f16(cross, umma_desc(a_h + ks * 32), umma_desc(b_l + ks * 32), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks)
          umma_f16(acc, umma_desc(a_h + ks * 32), umma_desc(b_h + ks * 32), idesc, ks > 0 ? 1u : 0u);
        umma_commit(&empty[st]);                 // the stage may be refilled
        umma_commit(&accfull[ab]);               // the slice's partial sums are complete
      }
    }
    __syncwarp();
  } else {
    for (int k = 0; k < nk; ++k) {
      const int ab = k & 1;
      mbar_wait_bounded(&accfull[ab], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      drain_accumulator((uint32_t)(ab * kAccStride), [](float &a, float v) { a += v; });
      tc_fence_before();                         // the drain's tcgen05.ld before the arrival / the next MMAs
      __syncwarp();
      if (lane == 0) mbar_arrive(&accfree[ab]);
    }
    // every MMA has completed (the last commit covers all earlier ones): add the cross terms, undo the 2^15 x 2^15
    drain_accumulator((uint32_t)kCrossCol, [](float &a, float v) { a = fmaf(v, 0x1p-11f, a) * 0x1p-30f; });

    // ---- accuracy guard.  A float16 operand carries its probability to 2^-22 relative only down to 2^-29 of the
    //      row maximum and to 2^-51 absolute below that (h underflows, the pre-scaled low term lives on), so the
    //      sum is off by at most 2^-51 (sum_c p_am + sum_c p_lm) absolutely: negligible unless Z itself is tiny -
    //      am and lm rows whose mass sits on DIFFERENT classes, tens of nats apart (the float32 reference keeps
    //      such a Z down to e^-87).  Cells with Z < 2^-27 (sum p_am + sum p_lm) (4 x the bound at 2^-22 relative)
    //      are recomputed exactly: the warp walks the two rows of the cell together, float32 products. ----
    {
      const float my_amneg = -sm.ammax[erow] * kLog2e;
      bool any = false;
#pragma unroll
      for (int i = 0; i < kColsPerHalf; ++i) {
        const int j = col_of(half, i);
        any |= t_ok && j < n_rows && s0 + j < S1 && accr[i] < (sa + sm.lmsum_g[j]) * 0x1p-27f;
      }
      if (__any_sync(0xffffffffu, any)) {        // never on ordinary data: one vote per warp is all the guard costs
#pragma unroll 1
      for (int i = 0; i < kColsPerHalf; ++i) {
        const int j = col_of(half, i), s = s0 + j;
        const bool cell = t_ok && j < n_rows && s < S1;
        unsigned need = __ballot_sync(0xffffffffu, cell && accr_at(i) < (sa + sm.lmsum_g[j]) * 0x1p-27f);
        while (need) {                           // warp-uniform
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int tt = t0 + q * 32 + src;
          const float amneg = __shfl_sync(0xffffffffu, my_amneg, src);
          const float lmneg = -sm.lmmax[j] * kLog2e;
          float acc = 0.f;
          for (int c = lane; c < C; c += 32) {
            float xa, xl;
            const size_t ia = ((size_t)b * p.T + tt) * C + c, il = ((size_t)b * S1 + s) * C + c;
            if (p.raw_dtype == FRN_F32) {
              xa = static_cast<const float *>(p.am_raw)[ia]; xl = static_cast<const float *>(p.lm_raw)[il];
            } else if (p.raw_dtype == FRN_BF16) {
              xa = __uint_as_float((uint32_t)static_cast<const unsigned short *>(p.am_raw)[ia] << 16);
              xl = __uint_as_float((uint32_t)static_cast<const unsigned short *>(p.lm_raw)[il] << 16);
            } else {
              xa = __half2float(static_cast<const __half *>(p.am_raw)[ia]);
              xl = __half2float(static_cast<const __half *>(p.lm_raw)[il]);
            }
            // exp(am - ammax) * exp(lm - lmmax) as ONE exponential: no intermediate underflow
            acc += ex2_approx(fmaf(xa, kLog2e, amneg) + fmaf(xl, kLog2e, lmneg));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (lane == src) accr_set(i, acc);
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();                               // every warp has left the operand stages: the epilogue reuses them
  FRN_TCT(1)
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
      if (epi && c0 < n_rows) {          // warp-uniform
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
          if (epi && c0 < n_rows) {             // warp-uniform
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
      if (pass == 0) { FRN_TCT(2) }
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
      FRN_TCT(3 + pass)
    }
  }
  tc_fence_before();
  __syncthreads();
  FRN_TCT(5)
  if (w == 0) tmem_dealloc(tmem_d, kTmemCols);
}

// ---------------------------------------------------------------------------
// host side
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

// One float16 operand plane [B][rows][Cp] as a 3-D tensor (column, row, utterance) that ENDS at column C: a box of
// 64 columns x box_rows rows lands in shared memory in the K-major SWIZZLE_128B layout tcgen05 reads; whatever
// lies beyond C, beyond the utterance's rows or beyond B arrives as zeros.
static bool make_map_3d(CUtensorMap *map, const unsigned short *base, int rows, int C, int Cp, int B, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)Cp * 2, (cuuint64_t)rows * Cp * 2};
  cuuint32_t box[3] = {(cuuint32_t)tc::KC, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<unsigned short *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool simple_logprobs_tc_applicable(const void *lm, const void *am, int C) {
  // the row-statistics kernel builds the operands with 128-bit loads (C % 4 == 0, 16-byte aligned bases); the
  // contraction itself needs the driver's tensor-map encoder
  return C % 4 == 0 && ((reinterpret_cast<uintptr_t>(am) | reinterpret_cast<uintptr_t>(lm)) & 15u) == 0 &&
         get_encode_fn() != nullptr;
}

// returns FRN_EUNSUPPORTED when the tensor-core path does not apply (caller falls back to the SIMT kernel):
// C % 4 != 0 or misaligned bases.  sp.XY != nullptr selects the arc-plane output (regular / modified only).
int launch_simple_logprobs_tc(const SimpleParams &sp, cudaStream_t stream) {
  if (sp.pxam_t == nullptr || sp.split.Cp == 0 || sp.gat.am_term == nullptr || get_encode_fn() == nullptr)
    return FRN_EUNSUPPORTED;                      // the row-statistics kernel did not prepare the operands
  if (sp.XY != nullptr && sp.rnnt_type == FRN_CONSTRAINED) return FRN_EUNSUPPORTED;
  CUtensorMap map_amh, map_aml, map_lmh, map_lml;
  if (!make_map_3d(&map_amh, sp.split.amh, sp.T, sp.C, sp.split.Cp, sp.B, tc::TM) ||
      !make_map_3d(&map_aml, sp.split.aml, sp.T, sp.C, sp.split.Cp, sp.B, tc::TM) ||
      !make_map_3d(&map_lmh, sp.split.lmh, sp.S + 1, sp.C, sp.split.Cp, sp.B, tc::TN) ||
      !make_map_3d(&map_lml, sp.split.lml, sp.S + 1, sp.C, sp.split.Cp, sp.B, tc::TN))
    return FRN_EUNSUPPORTED;
  auto kernel = sp.XY ? simple_logprobs_tc_kernel<true> : simple_logprobs_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
  if (e != cudaSuccess) return note_cuda_error(e);
  // arc-plane output: frames 0..T-1 carry arcs (the regular lattice's extra column T has none)
  dim3 grid(((sp.XY ? sp.T : sp.T1) + tc::TM - 1) / tc::TM, (sp.S + 1 + tc::TN - 1) / tc::TN, sp.B);
  count_launch(), kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(map_amh, map_aml, map_lmh, map_lml, sp);
  return check_launch();
}

}  // namespace frn

#ifdef FRN_TC_TIMING
extern "C" int frn_debug_tc_timing(long long *host_out) {
  return cudaMemcpyFromSymbol(host_out, frn::g_tc_timing, sizeof(frn::g_tc_timing)) == cudaSuccess ? 0 : 1;
}
#endif
