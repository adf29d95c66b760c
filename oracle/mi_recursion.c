/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * Plain-C restatement of the lattice recursion ("mutual information
 * recursion") that the reference runs on the GPU, its backward pass, and the
 * int32 cumulative minimum.  Sequential loops, one utterance after another;
 * the arithmetic follows the reference's equations, not its tiling.
 *
 * Reference being restated (paths under /root/reference):
 *   forward equation      tf_fast_rnnt/csrc/mutual_information_cuda.cu:141-152
 *   recursion spec        tf_fast_rnnt/python/tf_fast_rnnt/__init__.py:118-133
 *   LogAdd                tf_fast_rnnt/csrc/mutual_information.h:70-83
 *   boundary clipping     mutual_information_cuda.cu:264-279,295-303
 *   backward (3a)-(4b)    mutual_information_cuda.cu:472-481
 *   safe_exp              mutual_information_cuda.cu:430-439
 *   -1e30 clamp of p      mutual_information_cuda.cu:632-636
 *   grads zero outside    tf_fast_rnnt/python/csrc/tf_fast_rnnt_op.cc:94-98
 *   cummin (rhs>lhs)      mutual_information_cuda.cu:876-882
 *
 * Parity status: the reference's own tests hold no expected values
 * ("parity unpinned" by them).  This file is pinned instead against the
 * reference's CUDA kernels compiled from /root/reference into oracle/_ref and
 * run on the B200 (tests/test_gpu_ref_kernels.py).
 *
 * Built twice: -DREAL=float -DSUF=f32 and -DREAL=double -DSUF=f64.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifndef REAL
#define REAL float
#define SUF f32
#endif
#ifdef REAL_IS_DOUBLE
#define R_EXP exp
#define R_LOG1P log1p
#else
#define R_EXP expf
#define R_LOG1P log1pf
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

static inline REAL log_add(REAL x, REAL y) {
  /* mutual_information.h:70-83 */
  REAL diff;
  if (x < y) { diff = x - y; x = y; } else { diff = y - x; }
  if (diff - diff != 0) return x;
  return x + R_LOG1P(R_EXP(diff));
}

static inline REAL safe_exp(REAL x) {
  /* mutual_information_cuda.cu:430-439 */
  if (x - x != 0) return 0;
  REAL a = R_EXP(x);
  if (a - a != 0) return 0;
  return a;
}

/*
 * px: [B][S][T1]  (T1 == T+1 regular, T1 == T modified)
 * py: [B][S+1][T]
 * boundary: [B][4] = s_begin,t_begin,s_end,t_end
 * p:  [B][S+1][T+1] (out; cells outside the boundary box are left untouched)
 * ans: [B]
 */
void FN(orc_mi_forward)(const REAL *px, const REAL *py, const int32_t *boundary,
                        int B, int S, int T, int T1, REAL *p, REAL *ans) {
  const int modified = (T1 == T);
  const int off = modified ? -1 : 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const REAL *pxb = px + (size_t)b * S * T1;
    const REAL *pyb = py + (size_t)b * (S + 1) * T;
    REAL *pb = p + (size_t)b * (S + 1) * (T + 1);
    int s0 = boundary[4 * b + 0], t0 = boundary[4 * b + 1];
    int s1 = boundary[4 * b + 2], t1 = boundary[4 * b + 3];
    for (int s = s0; s <= s1; ++s) {
      for (int t = t0; t <= t1; ++t) {
        REAL v;
        if (s == s0 && t == t0) {
          v = 0;
        } else {
          REAL a = -INFINITY, c = -INFINITY;
          int tt = t + off;
          if (s > s0 && tt >= t0)
            a = pb[(size_t)(s - 1) * (T + 1) + tt] + pxb[(size_t)(s - 1) * T1 + tt];
          if (t > t0)
            c = pb[(size_t)s * (T + 1) + t - 1] + pyb[(size_t)s * T + t - 1];
          v = log_add(a, c);
        }
        pb[(size_t)s * (T + 1) + t] = v;
      }
    }
    ans[b] = (s1 >= s0 && t1 >= t0) ? pb[(size_t)s1 * (T + 1) + t1] : 0;
  }
}

/*
 * Backward with ans_grad (normally all ones).  p_grad: [B][S+1][T+1] scratch,
 * px_grad: [B][S][T1], py_grad: [B][S+1][T]; both are zero-filled here first.
 */
void FN(orc_mi_backward)(const REAL *px, const REAL *py, const int32_t *boundary,
                         const REAL *p, const REAL *ans_grad, int B, int S,
                         int T, int T1, REAL *p_grad, REAL *px_grad,
                         REAL *py_grad) {
  const int modified = (T1 == T);
  const int noff = modified ? 1 : 0;
  for (size_t i = 0; i < (size_t)B * S * T1; ++i) px_grad[i] = 0;
  for (size_t i = 0; i < (size_t)B * (S + 1) * T; ++i) py_grad[i] = 0;
  for (size_t i = 0; i < (size_t)B * (S + 1) * (T + 1); ++i) p_grad[i] = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const REAL *pxb = px + (size_t)b * S * T1;
    const REAL *pyb = py + (size_t)b * (S + 1) * T;
    const REAL *pb = p + (size_t)b * (S + 1) * (T + 1);
    REAL *gb = p_grad + (size_t)b * (S + 1) * (T + 1);
    REAL *gxb = px_grad + (size_t)b * S * T1;
    REAL *gyb = py_grad + (size_t)b * (S + 1) * T;
    int s0 = boundary[4 * b + 0], t0 = boundary[4 * b + 1];
    int s1 = boundary[4 * b + 2], t1 = boundary[4 * b + 3];
    if (s1 < s0 || t1 < t0) continue;
#define PCL(s, t) (pb[(size_t)(s) * (T + 1) + (t)] < (REAL)-1.0e30 ? (REAL)-1.0e30 : pb[(size_t)(s) * (T + 1) + (t)])
    for (int s = s1; s >= s0; --s) {
      for (int t = t1; t >= t0; --t) {
        REAL g;
        REAL gx = 0, gy = 0;
        /* arc (s,t) -> (s+1, t+noff) through px[s][t] */
        if (s < s1 && t + noff <= t1) {
          REAL term1 = safe_exp(PCL(s, t) + pxb[(size_t)s * T1 + t] - PCL(s + 1, t + noff));
          gx = gb[(size_t)(s + 1) * (T + 1) + t + noff] * term1;
        }
        /* arc (s,t) -> (s, t+1) through py[s][t] */
        if (t < t1) {
          REAL term2 = safe_exp(PCL(s, t) + pyb[(size_t)s * T + t] - PCL(s, t + 1));
          gy = gb[(size_t)s * (T + 1) + t + 1] * term2;
        }
        if (s == s1 && t == t1) g = ans_grad[b];
        else g = gx + gy;
        gb[(size_t)s * (T + 1) + t] = g;
        if (s < s1 && t + noff <= t1) gxb[(size_t)s * T1 + t] = gx;
        if (t < t1) gyb[(size_t)s * T + t] = gy;
      }
    }
#undef PCL
  }
}

#ifdef ORC_WITH_INT
/* inclusive running minimum along the last axis; mutual_information_cuda.cu:876-1012 */
void orc_cummin_i32(const int32_t *in, int32_t *out, int rows, int n) {
  for (int r = 0; r < rows; ++r) {
    int32_t m = INT32_MAX;
    for (int i = 0; i < n; ++i) {
      int32_t v = in[(size_t)r * n + i];
      if (m > v) m = v;
      out[(size_t)r * n + i] = m;
    }
  }
}
#endif
