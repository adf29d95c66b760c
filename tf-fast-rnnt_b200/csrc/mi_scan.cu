// Lattice recursion as a sequence of ROW SCANS (sm_100a).
//
// Same job as mi_dp.cu (the reference's K1/K2, tf_fast_rnnt/csrc/
// mutual_information_cuda.cu:174-422 forward, 490-760 backward), other
// parallelisation.  The wavefront needs S+T dependent steps.  Along a lattice
// row, however, the recursion (__init__.py:118-133, cu:141-152)
//
//     p[s][t] = p[s][t-1] * py[s][t-1]  +  p[s-1][t+off] * px[s-1][t+off]
//             =        y_t * p[s][t-1]  +  c_t
//
// is a first-order LINEAR recurrence in t whose inhomogeneous part c_t only
// needs the previous row: composing the affine maps  x -> y_t x + c_t  is
// associative, so a row is one parallel prefix scan and the critical path is
// (S+1) x (log2(32) + 1) compositions instead of S+T+1 steps.  A composition
// costs about as much as a wavefront step (shuffle, exponent alignment,
// multiply-add), so the scan pays for long lattices - T >= 8 S, or more than 512
// columns - and the wavefront kernel keeps the short ones (scan_dp_supported:
// measured table).  Threads own lattice COLUMNS, so px/py are read in the
// reference layout (t is the unit stride: coalesced), no skewed copy of the
// lattice is made, alpha/beta rows leave coalesced, and a ragged batch costs
// what its own rows and columns cost (warps beyond T_b exit at once).
//
// Numerics: the extended-range linear domain of mi_dp.cu (value = m * 2^e,
// float32 mantissa, exact int32 exponent; common.cuh), here with one exponent
// per CELL.  Composing two maps aligns the two terms of the new offset on the
// larger exponent with exact power-of-two factors; all terms are non-negative,
// so a term that falls out of range is negligible against the partial sum it is
// added to and nothing can under- or overflow.  No transcendental on the chain
// (arcs are decoded to (mantissa, exponent) one row ahead of their use).
//
// Execution: one chain of warps per (utterance, direction); a thread owns K = 1,
// 2 or 4 consecutive lattice columns, a warp 32 K.  Per row a warp folds its
// threads' columns, scans the 32 thread totals with 5 branch-free Kogge-Stone
// levels, takes the value of the column on its left from the previous warp of
// the chain, finishes its columns and posts its last one to the next warp.
// The hand-off goes through 8-deep mail boxes in shared memory (one 64-bit
// store posts {mantissa, exponent}, the reader empties the box; no block
// barrier), so warp w runs row s while warp w+1 runs row s-1 and the critical
// path is (S+1) warp scans plus one hand-off per warp.  For rows of more than
// 512 columns the chain is spread over a thread-block cluster of 2-8 CTAs (the
// mail box of a CTA's first warp is written through distributed shared memory)
// so that a thread still owns one column; with many chains (no room for
// clusters) a thread takes 2 or 4 columns instead.
// The backward recursion (cu:441-487) is the same scan on mirrored columns and
// descending rows.  Occupation counts are formed by scan_finalize_kernel from
// the alpha and beta planes (cu:472-481 in closed form: alpha * arc * beta' / Z).
#include "common.cuh"
#include "launchers.h"

namespace frn {

struct ScanParams {
  const float *px, *py;      // reference layout [B][S][T1], [B][S+1][T]
  const int32_t *boundary;   // [B][4]
  int2 *Aa, *Bb;             // [B][S+1][NP] {mantissa bits, exponent}, indexed by (s - s_begin, t - t_begin)
  int S, T, T1, NP;
  float delay_penalty;       // added to px (rnnt_loss.py:316-321); 0 = none
  float *ans, *px_grad, *py_grad;
  int cluster, wpc;          // CTAs per (utterance, direction); warps per CTA
};

constexpr int kScanRing = 8;   // mail boxes between two neighbouring warps
constexpr int kScanPF = 4;     // lattice rows of px/py in flight per thread (L2/DRAM latency ~ 3 row times)
constexpr int kScanMaxWarps = 16;
constexpr int kScanMaxK = 4;   // lattice columns per thread
constexpr int kBoxEmpty = INT32_MIN;

__device__ __forceinline__ bool scan_boundary_ok(const int4 &bd, int S, int T) {
  return bd.z - bd.x >= 0 && bd.w - bd.y >= 0 && bd.x >= 0 && bd.y >= 0 && bd.z <= S && bd.w <= T;
}
// natural-log score -> (mantissa in [1,2], exponent).  Scores below -2^16 in log2 units (-45 000 nats) are dead
// arcs (0, kNegI): with paths of at most 4096 arcs every live value then has an exponent above kNegI = -2^28, the
// exponent of a dead value, and no sum of two exponents leaves int32.
__device__ __forceinline__ void decode_arc(float v, float &m, int &e) {
  const float v2 = v * kLog2e;
  const float fl = floorf(v2);
  const bool alive = v2 > -65536.f;
  m = alive ? ex2_approx(v2 - fl) : 0.f;
  e = alive ? (int)fl : kNegI;
}
// (m, e) <- am * 2^ae * (bm * 2^be) + (m * 2^e), every operand >= 0
__device__ __forceinline__ void ext_fma(float am, int ae, float bm, int be, float &m, int &e) {
  const float tm = am * bm;
  const int te = ae + be;
  const int E = max(te, e);
  m = fmaf(tm, pow2i(te - E), m * pow2i(e - E));
  e = E;
}
// mantissa back into [1,2) (it is >= 1 or exactly 0 here); zero-safe and branch-free
__device__ __forceinline__ void renorm(float &m, int &e) {
  const int bits = __float_as_int(m);
  const int ex = max((bits >> 23) - 127, 0);
  e += ex;
  m = __int_as_float(bits - (ex << 23));
}

// mail box of the next warp of the chain: its state word, and posting a value into it.  The box may live
// in the next CTA of the cluster (distributed shared memory) unless the chain has a single CTA.
__device__ __forceinline__ int box_state(uint32_t addr, bool solo) {
  int v;
  if (solo) asm volatile("ld.volatile.shared.s32 %0, [%1+4];" : "=r"(v) : "r"(addr) : "memory");
  else asm volatile("ld.relaxed.cluster.shared::cluster.s32 %0, [%1+4];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void box_post(uint32_t addr, int m, int e, bool solo) {
  if (solo) asm volatile("st.volatile.shared.v2.s32 [%0], {%1,%2};" ::"r"(addr), "r"(m), "r"(e) : "memory");
  else asm volatile("st.relaxed.cluster.shared::cluster.v2.s32 [%0], {%1,%2};" ::"r"(addr), "r"(m), "r"(e) : "memory");
}

// SHIFT = 0: regular recursion (px has T+1 columns, vertical arcs); 1: modified (diagonal arcs).
// K lattice columns per thread.
template <int SHIFT, int K>
__device__ __forceinline__ void scan_dp_body(const ScanParams &p, int2 (*box)[kScanRing], const int4 bd, int b,
                                             int dir, int rank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = rank * p.wpc + w;               // position of this warp in the chain of the whole row
  const int s_begin = bd.x, t_begin = bd.y, Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const bool has_next = 32 * K * (gw + 1) <= Tb;
  const int j0 = (gw * 32 + lane) * K;           // scan positions j0 .. j0+K-1: t' = j forward, Tb - j backward
  const unsigned all = 0xffffffffu;

  // row r of the scan: forward s' = r, backward s' = Sb - r
  //   y_j = py[s'][t' - 1] (fwd)            | py[s'][t'] (bwd)        multiplies the left neighbour
  //   x_j = px[s' - 1][t' - SHIFT] (fwd)    | px[s'][t'] (bwd)        multiplies the previous row at j - SHIFT
  const int dj = dir == 0 ? 1 : -1;              // address step per scan position
  const int ystep = dir == 0 ? p.T : -p.T, xstep = dir == 0 ? p.T1 : -p.T1;
  const float *yp = p.py + ((size_t)b * (p.S + 1) + s_begin + (dir == 0 ? 0 : Sb)) * p.T + t_begin +
                    (dir == 0 ? j0 - 1 : Tb - j0);
  const float *xp = p.px + ((size_t)b * p.S + s_begin + (dir == 0 ? -1 : Sb)) * p.T1 + t_begin +
                    (dir == 0 ? j0 - SHIFT : Tb - j0);          // row of r = 0 (never read)
  int2 *plane = (dir == 0 ? p.Aa : p.Bb) + ((size_t)b * (p.S + 1) + (dir == 0 ? 0 : Sb)) * p.NP +
                (dir == 0 ? j0 : Tb - j0);
  const int pstep = dir == 0 ? p.NP : -p.NP;
  bool y_ok[K], x_ok[K], live[K];
  float pen[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    live[i] = j0 + i <= Tb;
    y_ok[i] = live[i] && j0 + i >= 1;
    x_ok[i] = live[i] && j0 + i >= SHIFT;
    const int xcol = t_begin + (dir == 0 ? j0 + i - SHIFT : Tb - j0 - i);
    pen[i] = p.delay_penalty != 0.f ? delay_penalty_value(bd.w, xcol, p.delay_penalty) : 0.f;
  }

  float rawy[kScanPF][K], rawx[kScanPF][K];
#pragma unroll
  for (int u = 0; u < kScanPF; ++u)
#pragma unroll
    for (int i = 0; i < K; ++i) {
      rawy[u][i] = rawx[u][i] = -INFINITY;
      if (y_ok[i] && u <= Sb) rawy[u][i] = __ldg(yp + (ptrdiff_t)u * ystep + i * dj);
      if (x_ok[i] && u >= 1 && u <= Sb) rawx[u][i] = __ldg(xp + (ptrdiff_t)u * xstep + i * dj);
    }
  yp += (ptrdiff_t)kScanPF * ystep;
  xp += (ptrdiff_t)kScanPF * xstep;
  // decoded arcs of the current row.  Row 0 has no row below it: it is seeded by a unit "previous row" at
  // position 0 (SHIFT = 1: by a unit carry into position 0) that goes through an identity arc.
  float ym[K], xm[K];
  int ye[K], xe[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    decode_arc(rawy[0][i], ym[i], ye[i]);
    xm[i] = 1.f;
    xe[i] = 0;
  }
  float vm[K], km = 0.f;         // previous row: own columns, and the column left of the warp (the carry it received)
  int ve[K], ke = kNegI;
#pragma unroll
  for (int i = 0; i < K; ++i) { vm[i] = 0.f; ve[i] = kNegI; }
  if (j0 == 0) {
    if (SHIFT) { km = 1.f; ke = 0; } else { vm[0] = 1.f; ve[0] = 0; }
  }
  // mail boxes: mine (local) and the next warp's, which may sit in the next CTA of the cluster
  const uint32_t box_in = smem_u32(&box[w][0]);
  uint32_t box_out;
  const bool solo = p.cluster == 1;               // one CTA per chain: plain shared-memory accesses
  {
    const bool local = w + 1 < p.wpc;
    const uint32_t laddr = smem_u32(&box[local ? w + 1 : 0][0]);
    const uint32_t target = local ? rank : (rank + 1 < p.cluster ? rank + 1 : rank);
    box_out = laddr;
    if (!solo) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(box_out) : "r"(laddr), "r"(target));
  }

  for (int r0 = 0; r0 <= Sb; r0 += kScanPF) {
#pragma unroll
    for (int u = 0; u < kScanPF; ++u) {
      const int r = r0 + u;
      if (r > Sb) break;
      const uint32_t slot = (uint32_t)(r % kScanRing) * 8u;
      // is the mail box of this row free again?  (asked early, needed at the end of the row)
      int out_state = kBoxEmpty;
      if (has_next) out_state = box_state(box_out + slot, solo);
      // inhomogeneous terms c_i = x_i * previous row at i - SHIFT
      float Bm[K], Am[K];
      int Be[K], Ae[K];
      if (SHIFT) {
        float sm = __shfl_up_sync(all, vm[K - 1], 1);
        int se = __shfl_up_sync(all, ve[K - 1], 1);
        if (lane == 0) { sm = km; se = ke; }
#pragma unroll
        for (int i = K - 1; i >= 1; --i) { Bm[i] = vm[i - 1] * xm[i]; Be[i] = ve[i - 1] + xe[i]; }
        Bm[0] = sm * xm[0];
        Be[0] = se + xe[0];
      } else {
#pragma unroll
        for (int i = 0; i < K; ++i) { Bm[i] = vm[i] * xm[i]; Be[i] = ve[i] + xe[i]; }
      }
      // local inclusive prefixes of the maps  x -> y_i x + c_i
      Am[0] = ym[0];
      Ae[0] = ye[0];
#pragma unroll
      for (int i = 1; i < K; ++i) {
        ext_fma(ym[i], ye[i], Bm[i - 1], Be[i - 1], Bm[i], Be[i]);
        Am[i] = Am[i - 1] * ym[i];
        Ae[i] = max(Ae[i - 1] + ye[i], kNegI);
      }
      // decode row r + 1, issue the loads of row r + PF into the slot just consumed
      float nym[K], nxm[K];
      int nye[K], nxe[K];
      {
        const bool more = r + kScanPF <= Sb;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          decode_arc(rawy[(u + 1) % kScanPF][i], nym[i], nye[i]);
          decode_arc(rawx[(u + 1) % kScanPF][i] + pen[i], nxm[i], nxe[i]);
          rawy[u][i] = rawx[u][i] = -INFINITY;      // untouched until decoded kScanPF - 1 rows later
          if (y_ok[i] && more) rawy[u][i] = __ldg(yp + i * dj);
          if (x_ok[i] && more) rawx[u][i] = __ldg(xp + i * dj);
        }
        yp += ystep;
        xp += xstep;
      }
      // inclusive scan of the threads' total maps over the warp (Kogge-Stone, branch-free)
      float SAm = Am[K - 1], SBm = Bm[K - 1];
      int SAe = Ae[K - 1], SBe = Be[K - 1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const bool up = lane >= o;
        float a1m = __shfl_up_sync(all, SAm, o);
        int a1e = __shfl_up_sync(all, SAe, o);
        float b1m = __shfl_up_sync(all, SBm, o);
        int b1e = __shfl_up_sync(all, SBe, o);
        a1m = up ? a1m : 1.f;
        a1e = up ? a1e : 0;
        b1m = up ? b1m : 0.f;
        b1e = up ? b1e : kNegI;
        ext_fma(SAm, SAe, b1m, b1e, SBm, SBe);
        SAm *= a1m;
        SAe = max(SAe + a1e, kNegI);
        // K columns per thread: a mantissa spanning 32 K columns would leave float32; one pull-back suffices
        if (K > 1 && o == 4) { renorm(SAm, SAe); renorm(SBm, SBe); }
      }
      // carry from the warp on the left: the value of the column left of this warp, in this row
      float cym = 0.f;
      int cye = kNegI;
      if (gw > 0) {
        int2 c;
        do {
          asm volatile("ld.volatile.shared.v2.s32 {%0,%1}, [%2];" : "=r"(c.x), "=r"(c.y) : "r"(box_in + slot));
        } while (c.y == kBoxEmpty);
        if (lane == 0)
          asm volatile("st.volatile.shared.v2.s32 [%0], {%1,%2};" ::"r"(box_in + slot), "r"(0), "r"(kBoxEmpty) : "memory");
        cym = __int_as_float(c.x);
        cye = c.y;
      }
      ext_fma(SAm, SAe, cym, cye, SBm, SBe);       // value of this thread's last column
      renorm(SBm, SBe);
      if (has_next) {
        while (out_state != kBoxEmpty) out_state = box_state(box_out + slot, solo);
        if (lane == 31) box_post(box_out + slot, __float_as_int(SBm), SBe, solo);
      }
      if (K > 1) {
        // value entering this thread from the left, then the other columns of the thread
        float im = __shfl_up_sync(all, SBm, 1);
        int ie = __shfl_up_sync(all, SBe, 1);
        if (lane == 0) { im = cym; ie = cye; }
#pragma unroll
        for (int i = 0; i < K - 1; ++i) {
          ext_fma(Am[i], Ae[i], im, ie, Bm[i], Be[i]);
          renorm(Bm[i], Be[i]);
        }
      }
      Bm[K - 1] = SBm;
      Be[K - 1] = SBe;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (live[i]) plane[i * dj] = make_int2(__float_as_int(Bm[i]), Be[i]);
        vm[i] = Bm[i]; ve[i] = Be[i];
        ym[i] = nym[i]; ye[i] = nye[i]; xm[i] = nxm[i]; xe[i] = nxe[i];
      }
      plane += pstep;
      km = cym; ke = cye;
    }
  }
}

template <int SHIFT, int K>
__global__ void __launch_bounds__(32 * kScanMaxWarps) scan_dp_kernel(ScanParams p) {
  __shared__ int2 box[kScanMaxWarps][kScanRing];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / p.cluster, dir = blockIdx.y;
  for (int i = threadIdx.x; i < kScanMaxWarps * kScanRing; i += blockDim.x) (&box[0][0])[i] = make_int2(0, kBoxEmpty);
  // every mail box of the cluster is initialised before anybody posts into it
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const int gw = (int)rank * p.wpc + (threadIdx.x >> 5);
  if (scan_boundary_ok(bd, p.S, p.T) && 32 * K * gw <= bd.w - bd.y)    // else: no column of this utterance here
    scan_dp_body<SHIFT, K>(p, box, bd, b, dir, (int)rank);
  // nobody leaves while a neighbour may still look into its mail boxes
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Occupation counts from the two planes; writes every element of px_grad / py_grad
// (zeros outside the boundary box: replaces the reference's memsets, op.cc:94-98).
template <int SHIFT>
__global__ void __launch_bounds__(256) scan_finalize_kernel(ScanParams p) {
  const int b = blockIdx.z, s = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
  const int4 bd = *reinterpret_cast<const int4 *>(p.boundary + 4 * b);
  const bool ok = scan_boundary_ok(bd, p.S, p.T);
  const int s_begin = bd.x, t_begin = bd.y, Sb = bd.z - bd.x, Tb = bd.w - bd.y;
  const int2 *Aa = p.Aa + (size_t)b * (p.S + 1) * p.NP;
  const int2 *Bb = p.Bb + (size_t)b * (p.S + 1) * p.NP;
  int2 tot = make_int2(0, 0);
  if (ok) tot = Aa[(size_t)Sb * p.NP + Tb];
  const float tot_m = __int_as_float(tot.x);
  const bool dead = !ok || !(tot_m > 0.f);
  if (s == 0 && t == 0)
    p.ans[b] = !ok ? 0.f : lattice_score(make_float2(tot_m, __int_as_float(tot.y)));
  if (p.px_grad == nullptr) return;
  const float inv_tot = dead ? 0.f : 1.0f / tot_m;
  const int sp = s - s_begin, tp = t - t_begin;
  float gx = 0.f, gy = 0.f;
  if (!dead && sp >= 0 && sp <= Sb && tp >= 0 && tp <= Tb) {
    const int2 a = Aa[(size_t)sp * p.NP + tp];
    const float am = __int_as_float(a.x) * inv_tot;
    const int ae = a.y - tot.y;
    // arc (s,t)->(s+1,t+SHIFT): cu:727-746; arc (s,t)->(s,t+1): cu:747-753
    if (sp < Sb && tp + SHIFT <= Tb) {
      const int2 q = Bb[(size_t)(sp + 1) * p.NP + tp + SHIFT];
      float v = p.px[((size_t)b * p.S + s) * p.T1 + t];
      if (p.delay_penalty != 0.f) v += delay_penalty_value(bd.w, t, p.delay_penalty);
      float xm;
      int xe;
      decode_arc(v, xm, xe);
      gx = (am * __int_as_float(q.x)) * xm * pow2i(min(ae + q.y + xe, 96));
    }
    if (tp < Tb) {
      const int2 q = Bb[(size_t)sp * p.NP + tp + 1];
      float ym;
      int ye;
      decode_arc(p.py[((size_t)b * (p.S + 1) + s) * p.T + t], ym, ye);
      gy = (am * __int_as_float(q.x)) * ym * pow2i(min(ae + q.y + ye, 96));
    }
  }
  if (s < p.S && t < p.T1) p.px_grad[((size_t)b * p.S + s) * p.T1 + t] = gx;
  if (t < p.T) p.py_grad[((size_t)b * (p.S + 1) + s) * p.T + t] = gy;
}

// ---------------------------------------------------------------------------
// CTAs per (utterance, direction): the chain of warps of one lattice row is spread over a thread-block
// cluster while that still fits the GPU in one wave (the row scan is instruction-issue bound: fewer warps
// per SM sub-partition = faster rows).  FRN_SCAN_CLUSTER overrides (experiments).
static int scan_cluster_size(int chains, int T) {
  const int sms = 148;     // B200; only steers a heuristic (how many CTAs a chain may spread over)
  // one column per thread is the fastest variant: take the smallest cluster that allows it (measured on
  // B200: at T = 500 clusters of 2 / 4 change nothing, 0.083 ms; at T = 1500 a cluster of 4 with one column
  // per thread runs 0.49 ms, one CTA with four columns per thread 0.57 ms)
  int c = 1;
  while (c < 8 && 32 * kScanMaxWarps * c < T + 1 && chains * c * 2 <= 2 * sms) c *= 2;
  if (const int v = debug_env_int("FRN_SCAN_CLUSTER", 0)) {
    if (v == 1 || v == 2 || v == 4 || v == 8) c = v;
  }
  return c;
}
static int scan_cols_per_thread(int T, int cluster) {
  // fewest columns per thread that fit the lattice row into the cluster; FRN_SCAN_K overrides (experiments)
  int k = 1;
  while (32 * kScanMaxWarps * cluster * k < T + 1) k *= 2;
  if (const int v = debug_env_int("FRN_SCAN_K", 0)) {
    if ((v == 1 || v == 2 || v == 4) && 32 * kScanMaxWarps * cluster * v >= T + 1) k = v;
  }
  return k;
}
// Which recursion kernel runs a dense lattice.  The row scan has (S+1) x 6 dependent compositions, the
// wavefront S+T+1 cheaper steps; measured on B200 (scripts/sweep_mi.py, ms per fwd+bwd call, scan / chain):
//   T=500:  S=20 .053/.055  S=50 .054/.064  S=100 .083/.084  S=200 .20/.16  S=400 .42/.38   T=200 S=100 .075/.050
//   T=1000: S=100 .13/.15   S=250 .32/.30   T=1500: S=100 .11/.17  S=400 .46/.68
// so: long lattices (T >= 8 S), and beyond 512 columns when T >= 4 S or the wavefront needs > 1 row per lane.
// FRN_DP_SCAN=1 / FRN_DP_CHAIN=1 force one of them in the debug-hooks build (tests and A/B runs).
bool scan_dp_feasible(int S, int T) {      // hard limits: one CTA must be able to hold a row
  return S >= 0 && T + 1 <= 32 * kScanMaxWarps * kScanMaxK && S + T <= 4095;
}
bool scan_dp_supported(int S, int T) {     // feasible AND the faster of the two recursion kernels at this shape
  if (!scan_dp_feasible(S, T)) return false;
  if (debug_env_int("FRN_DP_CHAIN", 0) == 1) return false;
  if (debug_env_int("FRN_DP_SCAN", 0) == 1) return true;
  if (T >= 8 * S) return true;
  return T + 1 > 512 && (T >= 4 * S || S + 1 > 256);
}
size_t scan_dp_workspace_bytes(int B, int S, int T) {
  const size_t plane = round_up_sz((size_t)B * (S + 1) * round_up(T + 1, 32) * sizeof(int2), 256);
  return 2 * plane;
}
int launch_scan_dp(const float *px, const float *py, const int32_t *boundary, int B, int S, int T, int T1,
                   float delay_penalty, bool want_grad, void *workspace, float *ans, float *px_grad,
                   float *py_grad, cudaStream_t stream) {
  const int NP = round_up(T + 1, 32);
  const size_t plane = round_up_sz((size_t)B * (S + 1) * NP * sizeof(int2), 256);
  ScanParams sp{px, py, boundary, reinterpret_cast<int2 *>(workspace),
                reinterpret_cast<int2 *>(static_cast<char *>(workspace) + plane), S, T, T1, NP, delay_penalty,
                ans, want_grad ? px_grad : nullptr, want_grad ? py_grad : nullptr};
  const int ndir = want_grad ? 2 : 1;
  const int cluster = scan_cluster_size(B * ndir, T);
  const int K = scan_cols_per_thread(T, cluster);
  const int warps = (T + 1 + 32 * K - 1) / (32 * K);
  sp.cluster = cluster;
  sp.wpc = (warps + cluster - 1) / cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * cluster, ndir);
  cfg.blockDim = dim3(32 * sp.wpc);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaSuccess;
  dim3 fgrid(want_grad ? (T + 1 + 255) / 256 : 1, want_grad ? S + 1 : 1, B);
  const bool mod = T1 == T;
#define FRN_LAUNCH_SCAN(K_)                                                                        \
  count_launch();                                                                                  \
  le = mod ? cudaLaunchKernelEx(&cfg, scan_dp_kernel<1, K_>, sp) : cudaLaunchKernelEx(&cfg, scan_dp_kernel<0, K_>, sp);
  if (K == 1) { FRN_LAUNCH_SCAN(1) } else if (K == 2) { FRN_LAUNCH_SCAN(2) } else { FRN_LAUNCH_SCAN(4) }
#undef FRN_LAUNCH_SCAN
  if (le != cudaSuccess) return note_cuda_error(le);
  if (mod) count_launch(), scan_finalize_kernel<1><<<fgrid, 256, 0, stream>>>(sp);
  else count_launch(), scan_finalize_kernel<0><<<fgrid, 256, 0, stream>>>(sp);
  return check_launch();
}

}  // namespace frn
