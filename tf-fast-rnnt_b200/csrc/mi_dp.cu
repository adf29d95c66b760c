// Lattice recursion ("mutual information recursion") for sm_100a.
//
// Replaces the reference's K1/K2 (tf_fast_rnnt/csrc/mutual_information_cuda.cu:
// 174-422 forward, 490-760 backward, launched 2*(S/32+T/32+1) times from host
// loops at :799-808 and :860-872) with three launches:
//
//   skew     px/py [B][S][T1] / [B][S+1][T]  ->  XY [B][d][s'] (diagonal-major,
//            boundary masks and delay penalty applied).  Every arc probability
//            is stored as (mantissa in [1,2], integer exponent), so the chain
//            never evaluates a transcendental.  Fully parallel, HBM/L2 bound.
//   chain    one CTA per (utterance, direction), a software pipeline of warps
//            (see dp_chain_kernel); one step = one anti-diagonal (d = t' + k s',
//            k = 1 regular / 0 modified), neighbours exchanged by warp shuffle, the
//            diagonal-major arcs streamed into a shared-memory ring by 1-D bulk
//            async copies (TMA engine) behind mbarriers.  Forward (alpha) and
//            backward (beta) chains run concurrently in different CTAs.
//            The recursion runs in the LINEAR domain with an extended exponent:
//            value = m * 2^o (float32 mantissa, exact int32 frame per row), one
//            dependent shuffle + FMA per step instead of a log-add (shuffle +
//            ex2 + lg2); see the comment above ChainParams.
//            Latency bound: (S_b + T_b) dependent multiply-adds.
//   finalize occupation counts  px_grad = alpha * (px * beta') / total,
//            py_grad likewise (equal to the reference's p_grad recursion,
//            cu:472-481, in exact arithmetic), written in the reference layout.
//
#include "common.cuh"

namespace frn {

// ---------------------------------------------------------------------------
// skew (dense input)
// ---------------------------------------------------------------------------
struct SkewDenseParams {
  const float *px, *py;     // reference layout
  const int32_t *boundary;  // [B][4]
  float4 *XY;               // [B][Dn][P]
  int S, T, T1, P, Dn;
  float delay_penalty;      // added to px (rnnt_loss.py:316-321); 0 = none
};

// One block: 128 diagonals x 32 rows of one utterance (the kernel is instruction-issue bound, so
// the tile is wide: 1.24x over-read of px/py instead of 2x, no divisions in the index math).
constexpr int kSkewDiags = 128;
template <int K>
__global__ void __launch_bounds__(256) skew_dense_kernel(SkewDenseParams p) {
  constexpr int kWidth = kSkewDiags + 31 * K;  // t' values touched by the tile
  constexpr int kPitch = kWidth + (K ? 3 : 1);  // (kPitch - K) % 32 == 1: conflict-free transposed reads
  __shared__ float sx[32 * kPitch], sy[32 * kPitch];
  const int b = blockIdx.z;
  const int d0 = blockIdx.x * kSkewDiags, s0 = blockIdx.y * 32;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y, Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const int off = K ? 0 : -1;
  const int tlo = d0 - K * (s0 + 31);  // smallest t' of the tile
  const float *pxb = p.px + (size_t)b * p.S * p.T1;
  const float *pyb = p.py + (size_t)b * (p.S + 1) * p.T;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;

  // load: one warp per row s', lanes along t' (coalesced along t).  All loads of the thread (4 rows x 5
  // column chunks x 2 arrays) are issued before the first one is used: one memory round trip instead of 20.
  constexpr int kChunks = (kWidth + 31) / 32;
  float rx[4][kChunks], ry[4][kChunks];
#pragma unroll
  for (int ri = 0; ri < 4; ++ri) {
    const int row = wrp + 8 * ri;
    const int sp = s0 + row;  // destination row s'
    const bool row_ok = sp <= Sb;
    const float *pxr = pxb + (size_t)(s_begin + sp - 1) * p.T1 + t_begin + off;
    const float *pyr = pyb + (size_t)(s_begin + sp) * p.T + t_begin - 1;
#pragma unroll
    for (int ci = 0; ci < kChunks; ++ci) {
      const int col = ci * 32 + lane;
      const int tp = tlo + col;  // destination cell (s', t')
      rx[ri][ci] = ry[ri][ci] = kNeg;
      if (col < kWidth && row_ok && tp <= Tb) {
        // mutual_information_cuda.cu:295-303 (forward load rules)
        if (sp >= 1 && tp + off >= 0) rx[ri][ci] = __ldg(pxr + tp);
        if (tp >= 1) ry[ri][ci] = __ldg(pyr + tp);
      }
    }
  }
#pragma unroll
  for (int ri = 0; ri < 4; ++ri) {
    const int row = wrp + 8 * ri;
#pragma unroll
    for (int ci = 0; ci < kChunks; ++ci) {
      const int col = ci * 32 + lane;
      const int tp = tlo + col;
      float vx = rx[ri][ci], vy = ry[ri][ci];
      if (vx != kNeg) {      // a loaded arc (kNeg itself is dead anyway)
        if (p.delay_penalty != 0.f) vx += delay_penalty_value(bd.w, t_begin + tp + off, p.delay_penalty);
        vx = fmaxf(vx * kLog2e, kNeg);
      }
      if (vy != kNeg) vy = fmaxf(vy * kLog2e, kNeg);
      if (col < kWidth) {
        sx[row * kPitch + col] = vx;
        sy[row * kPitch + col] = vy;
      }
    }
  }
  __syncthreads();
  // store: one warp per diagonal, lanes along s' (one 16-byte arc pair per thread, 512 contiguous bytes per warp)
  float4 *XYb = p.XY + (size_t)b * p.Dn * p.P + s0 + lane;
  const int colbase = K * (31 - lane) + lane * kPitch;
  for (int dd = wrp; dd < kSkewDiags; dd += 8) {
    const int d = d0 + dd;
    if (d >= p.Dn) break;
    const float2 ax = encode_arc(sx[colbase + dd]), ay = encode_arc(sy[colbase + dd]);
    XYb[(size_t)d * p.P] = make_float4(ax.x, ax.y, ay.x, ay.y);
  }
}

// ---------------------------------------------------------------------------
// chain
// ---------------------------------------------------------------------------
// Numerics.  Every lattice value is carried as  m * 2^o : `o` an exact int32
// frame, `m` a float32 mantissa (0 = dead).  One step of a row is
//     EA = o_feeder + e_x,  EB = o_own + e_y            (exact integer frames of the two terms)
//     o' = max(EA, EB)
//     m' = m_feeder * (x_m * 2^(EA-o')) + m_own * (y_m * 2^(EB-o'))
// Every re-scaling is by an exact power of two built with integer ops, the step
// has no transcendental and its only roundings are the products and the FMA.
// The frames depend on integers only, so a step consists of two short independent
// dependency chains: (shuffle, add, max) for the frame and (shuffle, FMA) for the
// mantissa.  The dominant term keeps its mantissa (times an arc mantissa in
// [1,2]), so m never shrinks below 1 and grows by at most x4 per step; it is
// re-normalised to [1,2) (a bit operation, exponent moved into the frame) at
// every chunk boundary (<= 16 steps) and whenever it is handed to another warp.
// Nothing can under- or overflow whatever the arc scores are.  Against float64
// the occupation counts come out ~1e-6 relative at the c2 shape; the reference's
// plain float32 log-domain p[] is at 1.3e-3 there ("numerics").
struct ChainParams {
  const float4 *XY;
  float2 *A;           // [B][Dn][P]  forward {mantissa, frame} (dir 0)
  float4 *Bq;          // [B][Dn][P]  backward-side operands and their frame (dir 1)
  const int32_t *boundary;
  int k, P, Dn, S, T;
  int CH, NST;         // diagonals per bulk copy; ring stages (> warps)
};

#ifdef FRN_CHAIN_TIMING
// diagnostic build only (scripts/chain_timing.py): cycles of block 0's warps spent waiting for arcs, waiting for
// the feeding warp, in the steps, and in total: [direction][warp][4]
__device__ unsigned long long g_chain_timing[2][8][4];
#define FRN_CT(x) x
#else
#define FRN_CT(x)
#endif

// Execution.  One CTA per (utterance, direction); RPL consecutive lattice rows
// per lane, 32*RPL rows per warp, W = P / (32*RPL) warps.  The warps form a
// software pipeline: warp w runs one chunk (CH diagonals) behind the warp that
// owns the rows feeding it and picks the boundary row's (mantissa, frame) of
// every step out of a shared-memory ring that warp filled; one mbarrier per
// (warp, ring slot) says "chunk published".  There is no per-step block barrier
// and every warp sits alone on its SM sub-partition.  XY chunks are shared by
// all warps: NST > W stages of 1-D bulk copies; the tail warp of the pipeline
// recycles a stage.
//
// PC != 0: the plane pitch P is the compile-time constant PC and a chunk is kChunk diagonals.  A full chunk
// then runs as one straight-line block of kChunk steps in which every shared- and global-memory address is a
// constant offset from the chunk's base pointers (no per-step pointer arithmetic): the step is bound by
// instruction issue on one SM sub-partition, and this form issues about half the instructions per step.
template <int RPL, int DIR, int PC>
__device__ __forceinline__ void dp_chain_body(const ChainParams &p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5;
  const int P = PC ? PC : p.P, CH = PC ? kChunk : p.CH, NST = p.NST;
  const int stage_elems = CH * P;
  // one diagonal of padding on either side of the ring: the operand prefetch of the last step of a
  // chunk reads one diagonal past it
  float4 *ring = reinterpret_cast<float4 *>(smem_raw) + P;
  uint64_t *mbar_xy = reinterpret_cast<uint64_t *>(ring + (size_t)NST * stage_elems + P);
  uint64_t *mbar_edge = mbar_xy + NST;                               // [W][NST], indexed by consumer warp
  float2 *edge = reinterpret_cast<float2 *>(mbar_edge + W * NST);    // [W][NST][CH], indexed by consumer warp
  float2 *dead_edge = edge + (size_t)W * NST * CH;                   // CH entries: what an unfed warp reads

  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  if (Sb < 0 || Tb < 0 || bd.x < 0 || bd.y < 0 || bd.z > p.S || bd.w > p.T) return;  // finalize reports it
  const int Db = Tb + p.k * Sb;
  const int nchunk = Db / CH + 1;
  const size_t plane = (size_t)b * p.Dn * P;
  const float4 *XYg = p.XY + plane;
  const uint32_t chunk_bytes = (uint32_t)(stage_elems * sizeof(float4));

  auto issue = [&](int seq, int st) {
    const int ci = DIR ? nchunk - 1 - seq : seq;
    mbar_arrive_expect_tx(&mbar_xy[st], chunk_bytes);
    bulk_g2s(ring + (size_t)st * stage_elems, XYg + (size_t)ci * stage_elems, chunk_bytes, &mbar_xy[st]);
  };

  if (tid == 0) {
    for (int i = 0; i < NST + W * NST; ++i) mbar_init(&mbar_xy[i], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < CH; i += blockDim.x) dead_edge[i] = make_float2(0.f, __int_as_float(kNegI));
  // Backward, last lane of the lattice: the symbol arc "into row r0 + RPL" is read at row P of a diagonal,
  // i.e. at row 0 of the next diagonal in the ring.  Row 0 has no incoming symbol arc, so that entry is a dead
  // arc in every diagonal the bulk copies deliver - but behind the last diagonal of the last stage lies the
  // padding, and a stage whose first copy is still in flight holds whatever the shared memory held before: a
  // garbage exponent there inflated the frames of the top rows and flushed their mantissas (found by the
  // scan-vs-wavefront fuzz test).  Row 0 of every ring diagonal and of the padding starts out dead.
  for (int i = tid; i <= NST * CH; i += blockDim.x)
    ring[(size_t)i * P] = make_float4(0.f, __int_as_float(kNegI), 0.f, __int_as_float(kNegI));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ordered before the bulk copies into the ring
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < NST && i < nchunk; ++i) issue(i, i);

  const int pos = DIR ? (W - 1 - w) : w;        // position in the warp pipeline, 0 = head
  const bool is_tail = (pos == W - 1), fed = (pos > 0), feeds = (pos < W - 1);
  const int wc = DIR ? w - 1 : w + 1;           // the warp this one feeds
  const int r0 = RPL * tid;                     // first lattice row of this lane
  const bool lane_in = DIR ? (lane == 31) : (lane == 0);    // lane fed across the warp boundary
  const bool lane_out = DIR ? (lane == 0) : (lane == 31);   // lane feeding the next warp
  const bool publish = lane_out && feeds;
  constexpr int step_sign = DIR ? -1 : 1;
  const float2 dead2 = make_float2(0.f, __int_as_float(kNegI));
  float m[RPL];
  int o[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const bool src = (r0 + j) == (DIR ? Sb : 0);
    m[j] = src ? 1.f : 0.f;
    o[j] = src ? 0 : kNegI;
  }
  float2 *const outA = p.A + plane + r0;
  float4 *const outB = p.Bq + plane + r0;
  if (!DIR) {
#pragma unroll
    for (int j = 0; j < RPL; ++j) outA[j] = make_float2(m[j], __int_as_float(o[j]));
  }
  // state of the row feeding lane_in before the first step (initial condition)
  float2 carry = dead2;
  if (DIR && fed && (r0 + RPL) == Sb) carry = make_float2(1.f, 0.f);

  // Per-chunk state is carried incrementally (no divisions, no 64-bit re-derivations per chunk).
  int st = 0;                                   // ring stage of chunk i
  uint32_t par = 0;                             // its mbarrier phase parity
  int cbase = DIR ? (nchunk - 1) * CH : 0;      // first diagonal of chunk i
  const float4 *xstage = ring + r0;             // this lane's column of stage `st`
  const float2 *estage_in = fed ? edge + (size_t)w * NST * CH : dead_edge;     // feeding row's states, stage `st`
  float2 *estage_out = edge + (size_t)(feeds ? wc : w) * NST * CH;            // our states for the fed warp
  uint64_t *bar_xy = mbar_xy, *bar_in = mbar_edge + w * NST, *bar_out = mbar_edge + wc * NST;
  // outputs walk the diagonals contiguously across chunks
  float2 *pa = outA + (size_t)(DIR ? min(cbase + CH - 1, Db) : 1) * P;
  float4 *pb = outB + (size_t)((DIR ? min(cbase + CH - 1, Db) : 1) - 1) * P;
  const float4 *gnext = XYg + (size_t)(DIR ? (nchunk - 1 - NST) : NST) * stage_elems;   // next chunk to request

  FRN_CT(long long ct_xy = 0; long long ct_in = 0; long long ct_steps = 0; const long long ct_begin = clock64();)
  for (int i = 0; i < nchunk; ++i) {
    FRN_CT(const long long ct0 = clock64();)
    mbar_wait(bar_xy, par);
    FRN_CT(const long long ct1 = clock64();)
    if (fed) mbar_wait(bar_in, par);            // the feeding warp has published this chunk
    FRN_CT(const long long ct2 = clock64(); ct_xy += ct1 - ct0; ct_in += ct2 - ct1;)
    const int e_lo = max(cbase, 1), e_hi = min(cbase + CH - 1, Db);
    const int n = e_hi - e_lo + 1;
    const int el0 = (DIR ? e_hi : e_lo) - cbase;
    // running pointers, all advanced by one diagonal per step
    const float4 *xp = xstage + el0 * P;                                 // arcs of the current step
    const float2 *pe = estage_in + el0;                                  // feeding row's state
    float2 *po = estage_out + el0;                                       // our state for the fed warp
    // operands of the first step of the chunk
    float4 a4[RPL];
    float2 xnext = dead2;
    // The feeding warp publishes its values as they are (un-normalised, like the ones the lanes of a warp hand
    // each other): both warps were normalised at the same chunk boundary, so after q steps both are bounded by
    // 2 * 4^q and the bound does not compound.  Only the value carried over a chunk boundary is normalised here.
    float2 ev = carry;
    {
      int ce = __float_as_int(ev.y);
      normalise_pair(ev.x, ce);
      ev.y = __int_as_float(ce);
    }
#pragma unroll
    for (int j = 0; j < RPL; ++j) a4[j] = make_float4(0.f, dead2.y, 0.f, dead2.y);
    if (n > 0) {
#pragma unroll
      for (int j = 0; j < RPL; ++j) a4[j] = xp[j];
      if (DIR) xnext = *reinterpret_cast<const float2 *>(xp + RPL);
    }
    int nb_o_sh = DIR ? __shfl_down_sync(0xffffffffu, o[0], 1) : __shfl_up_sync(0xffffffffu, o[RPL - 1], 1);
    // One step; `so` = its diagonal relative to the running pointers (0 in the generic loop, which advances
    // the pointers every step; a compile-time constant in the straight-line form of a full chunk).
    auto do_step = [&](const int so) {
      // ---- prefetch the operands of the next step (off the dependency chain).  After the last
      // step of a chunk this reads one diagonal past the chunk (the ring is padded); unused. ----
      float4 b4[RPL];
      float2 xnext2 = dead2;
#pragma unroll
      for (int j = 0; j < RPL; ++j) b4[j] = xp[(so + step_sign) * P + j];
      if (DIR) xnext2 = *reinterpret_cast<const float2 *>(xp + (so + step_sign) * P + RPL);
      const float2 ev2 = pe[so];   // the feeding row's state after ITS step e = input of our next step (broadcast read)

      // ---- neighbour across the lane boundary (its frame was shuffled as soon as it was known) ----
      float nb_m = DIR ? __shfl_down_sync(0xffffffffu, m[0], 1) : __shfl_up_sync(0xffffffffu, m[RPL - 1], 1);
      nb_m = lane_in ? ev.x : nb_m;
      const int nb_o = lane_in ? __float_as_int(ev.y) : nb_o_sh;

      float raw[RPL];
      int on[RPL];
      if (!DIR) {
        // alpha_e(s') = alpha_{e-1}(s'-1) * X[e][s'] + alpha_{e-1}(s') * Y[e][s']
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const float fm = j ? m[j ? j - 1 : 0] : nb_m;
          const int fo = j ? o[j ? j - 1 : 0] : nb_o;
          const int EA = fo + __float_as_int(a4[j].y), EB = o[j] + __float_as_int(a4[j].w);
          on[j] = max(max(EA, EB), kNegI);
          if (j == RPL - 1) nb_o_sh = __shfl_up_sync(0xffffffffu, on[j], 1);     // frame chain runs ahead
          const float gx = a4[j].x * pow2i(EA - on[j]);
          raw[j] = fmaf(fm, gx, m[j] * (a4[j].z * pow2i(EB - on[j])));
          pa[so * P + j] = make_float2(raw[j], __int_as_float(on[j]));
        }
      } else {
        // beta_{e-1}(s') = X[e][s'+1] * beta_e(s'+1) + Y[e][s'] * beta_e(s')
#pragma unroll
        for (int j = RPL - 1; j >= 0; --j) {
          const bool last = (j == RPL - 1);
          const float fm = last ? nb_m : m[last ? j : j + 1];
          const int fo = last ? nb_o : o[last ? j : j + 1];
          const float xm = last ? xnext.x : a4[last ? j : j + 1].x;
          const int xe = __float_as_int(last ? xnext.y : a4[last ? j : j + 1].y);
          const int EA = fo + xe, EB = o[j] + __float_as_int(a4[j].w);
          on[j] = max(max(EA, EB), kNegI);
          if (j == 0) nb_o_sh = __shfl_down_sync(0xffffffffu, on[j], 1);
          const float gx = xm * pow2i(EA - on[j]);
          const float c = m[j] * (a4[j].z * pow2i(EB - on[j]));
          raw[j] = fmaf(fm, gx, c);
          // operands of diagonal e-1, expressed in the frame `on`
          pb[so * P + j] = make_float4(fm * gx, c, __int_as_float(on[j]), 0.f);
        }
      }
      // The frames evolve by integer ops only and the mantissas by one FMA: two short, independent
      // dependency chains per step.  Mantissas are left un-normalised inside a chunk (they grow by
      // at most x4 per step) and are re-normalised at the chunk boundary.
#pragma unroll
      for (int j = 0; j < RPL; ++j) { m[j] = raw[j]; o[j] = on[j]; }
      if (publish) po[so] = make_float2(DIR ? m[0] : m[RPL - 1], __int_as_float(DIR ? o[0] : o[RPL - 1]));
      // rotate the prefetched operands in
#pragma unroll
      for (int j = 0; j < RPL; ++j) a4[j] = b4[j];
      xnext = xnext2;
      ev = ev2;
    };
    if (PC != 0 && n == kChunk) {
#pragma unroll
      for (int q = 0; q < kChunk; ++q) do_step(q * step_sign);
      xp += step_sign * kChunk * P;
      pe += step_sign * kChunk;
      po += step_sign * kChunk;
      pa += step_sign * kChunk * P;
      pb += step_sign * kChunk * P;
    } else {
#pragma unroll 2
      for (int q = 0; q < n; ++q) {
        do_step(0);
        xp += step_sign * P;
        pe += step_sign;
        po += step_sign;
        pa += step_sign * P;
        pb += step_sign * P;
      }
    }
    FRN_CT(ct_steps += clock64() - ct2;)
#pragma unroll
    for (int j = 0; j < RPL; ++j) normalise_pair(m[j], o[j]);
    carry = ev;   // the feeding row's state after the last step of this chunk
    if (publish) mbar_arrive(bar_out);
    if (is_tail) {
      __syncwarp();
      if (lane == 0 && i + NST < nchunk) {      // recycle the stage: request chunk i + NST into it
        mbar_arrive_expect_tx(bar_xy, chunk_bytes);
        bulk_g2s(ring + (size_t)st * stage_elems, gnext, chunk_bytes, bar_xy);
      }
    }
    gnext += step_sign * stage_elems;
    cbase += step_sign * CH;
    ++st; ++bar_xy; ++bar_in; ++bar_out;
    xstage += stage_elems; estage_out += CH;
    if (fed) estage_in += CH;
    if (st == NST) {
      st = 0; par ^= 1u;
      bar_xy -= NST; bar_in -= NST; bar_out -= NST;
      xstage -= (size_t)NST * stage_elems; estage_out -= NST * CH;
      if (fed) estage_in -= NST * CH;
    }
  }
#ifdef FRN_CHAIN_TIMING
  if (b == 0 && lane == 0 && w < 8) {
    unsigned long long *t = g_chain_timing[DIR][w];
    t[0] = ct_xy; t[1] = ct_in; t[2] = ct_steps; t[3] = clock64() - ct_begin;
  }
#endif
}

// grid = (B, 2): blockIdx.y selects the direction so that the forward and the
// backward chain of every utterance run concurrently on different SMs.
template <int RPL, int PC>
__global__ void __launch_bounds__(256, 1) dp_chain_kernel(ChainParams p) {
  if (blockIdx.y == 0) dp_chain_body<RPL, 0, PC>(p);
  else dp_chain_body<RPL, 1, PC>(p);
}

// ---------------------------------------------------------------------------
// finalize (dense output)
// ---------------------------------------------------------------------------
struct FinalizeDenseParams {
  const float2 *A;
  const float4 *Bq;
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
  p.ans[b] = lattice_score(p.A[((size_t)b * p.Dn + Tb + p.k * Sb) * p.P + Sb]);
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
  // tiles of the skewed plane that hold no lattice cell at all (t = d - K s outside [0, T] for every row of
  // the tile: the corners of the parallelogram) have nothing to read and nothing to write
  if (K && (d0 + 31 < s0 || d0 - min(s0 + 31, p.S) > p.T) && !(blockIdx.x == 0 && blockIdx.y == 0)) return;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int s_begin = bd.x, t_begin = bd.y;
  const int Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const bool ok = boundary_ok(bd, p.S, p.T);
  const int noff = K ? 0 : 1;
  const int Db = Tb + K * Sb;
  const size_t plane = (size_t)b * p.Dn * p.P;
  float2 tot = make_float2(0.f, 0.f);
  if (ok) tot = p.A[plane + (size_t)Db * p.P + Sb];
  const bool dead = !ok || !(tot.x > 0.f);
  const float inv_tot = dead ? 0.f : 1.0f / tot.x;
  const int tot_o = __float_as_int(tot.y);
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
          const float2 a = p.A[at];
          const float4 bq = p.Bq[at];
          const float sc = occupation_scale(a, __float_as_int(bq.z), tot_o, inv_tot);
          // arc (s,t)->(s+1,t+noff): cu:727-746; arc (s,t)->(s,t+1): cu:747-753
          if (sp < Sb && tp + noff <= Tb) vx = bq.x * sc;
          if (tp < Tb) vy = bq.y * sc;
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
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.ans[b] = !ok ? 0.f : lattice_score(tot);
}

// ---------------------------------------------------------------------------
// host-side launchers (used by api.cu)
// ---------------------------------------------------------------------------
int launch_skew_dense(const float *px, const float *py, const int32_t *boundary, const DpGeom &g,
                      const DpWorkspace &w, float delay_penalty, cudaStream_t stream) {
  SkewDenseParams sp{px, py, boundary, w.XY, g.S, g.T, g.T1, g.P, g.Dn, delay_penalty};
  dim3 grid((g.Dn + kSkewDiags - 1) / kSkewDiags, g.P / 32, g.B);
  if (g.k) count_launch(), skew_dense_kernel<1><<<grid, 256, 0, stream>>>(sp);
  else count_launch(), skew_dense_kernel<0><<<grid, 256, 0, stream>>>(sp);
  return check_launch();
}

struct ChainConfig { int W, CH, NST; size_t smem; };
static ChainConfig chain_config(const DpGeom &g) {
  ChainConfig c;
  c.W = g.P / (32 * g.rpl);
  auto bytes = [&](int nst, int ch) {
    return (size_t)(nst * ch + 2) * g.P * sizeof(float4) + (size_t)(nst + c.W * nst) * sizeof(uint64_t) +
           (size_t)(c.W * nst * ch + ch) * sizeof(float2) + 128;
  };
  const size_t budget = 200 * 1024;
  // ring = one stage per warp of the pipeline + look-ahead; prefer 32 diagonals of look-ahead (a bulk
  // copy takes ~1.5k cycles, a step ~50) and long chunks (fewer mbarrier round trips)
  for (int la = 32; la >= 8; la >>= 1)
    for (int ch = kChunk; ch >= 4; ch >>= 1) {
      const int nst = c.W + max(2, (la + ch - 1) / ch);
      if (bytes(nst, ch) <= budget) { c.CH = ch; c.NST = nst; c.smem = bytes(nst, ch); return c; }
    }
  for (int ch = 2; ch >= 1; --ch)
    for (int extra = 16; extra >= 2; --extra)
      if (bytes(c.W + extra, ch) <= budget) { c.CH = ch; c.NST = c.W + extra; c.smem = bytes(c.NST, ch); return c; }
  c.CH = 1; c.NST = c.W + 2; c.smem = bytes(c.NST, 1);
  return c;
}

int launch_chain(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, bool both_directions,
                 cudaStream_t stream) {
  if (g.P > kMaxRowsDp) return FRN_EUNSUPPORTED;
  const ChainConfig c = chain_config(g);
  ChainParams cp{w.XY, w.A, w.Bq, boundary, g.k, g.P, g.Dn, g.S, g.T, c.CH, c.NST};
  dim3 grid(g.B, both_directions ? 2 : 1);
  const int threads = 32 * c.W;
  cudaError_t e;
#define FRN_LAUNCH_CHAIN(RPL_, PC_)                                                                                \
  e = cudaFuncSetAttribute(dp_chain_kernel<RPL_, PC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem); \
  if (e != cudaSuccess) return note_cuda_error(e);                                                                \
  count_launch(), dp_chain_kernel<RPL_, PC_><<<grid, threads, c.smem, stream>>>(cp);
  // straight-line chunks for the pitches of short label sequences (S + 1 <= 128: c1, c2, c3)
  const bool fixed = g.rpl == 1 && c.CH == kChunk && debug_env_int("FRN_CHAIN_GENERIC", 0) == 0;
  if (fixed && g.P == 128) { FRN_LAUNCH_CHAIN(1, 128) }
  else if (fixed && g.P == 96) { FRN_LAUNCH_CHAIN(1, 96) }
  else if (fixed && g.P == 64) { FRN_LAUNCH_CHAIN(1, 64) }
  else if (fixed && g.P == 32) { FRN_LAUNCH_CHAIN(1, 32) }
  else if (g.rpl == 1) { FRN_LAUNCH_CHAIN(1, 0) }
  else if (g.rpl == 2) { FRN_LAUNCH_CHAIN(2, 0) }
  else { FRN_LAUNCH_CHAIN(4, 0) }
#undef FRN_LAUNCH_CHAIN
  return check_launch();
}

int launch_finalize_dense(const int32_t *boundary, const DpGeom &g, const DpWorkspace &w, float *ans,
                          float *px_grad, float *py_grad, cudaStream_t stream) {
  FinalizeDenseParams fp{w.A, w.Bq, boundary, ans, px_grad, py_grad, g.S, g.T, g.T1, g.P, g.Dn, g.k};
  if (px_grad == nullptr || py_grad == nullptr) {
    count_launch(), dp_ans_kernel<<<(g.B + 127) / 128, 128, 0, stream>>>(fp, g.B);
    return check_launch();
  }
  // absolute diagonals dabs = t + k*s for t in [0, T], s in [0, S]
  dim3 grid((g.T + 1 + g.k * g.S + 31) / 32, (g.S + 1 + 31) / 32, g.B);
  if (g.k) count_launch(), finalize_dense_kernel<1><<<grid, 256, 0, stream>>>(fp);
  else count_launch(), finalize_dense_kernel<0><<<grid, 256, 0, stream>>>(fp);
  return check_launch();
}

}  // namespace frn

#ifdef FRN_CHAIN_TIMING
extern "C" int frn_debug_chain_timing(unsigned long long *host_out) {
  return cudaMemcpyFromSymbol(host_out, frn::g_chain_timing, sizeof(frn::g_chain_timing)) == cudaSuccess ? 0 : 1;
}
#endif
