/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the CPU kernels behind k2's
 * `mutual_information_recursion`, which is what the reference reaches through
 *   /root/reference/model/joiner/joiner.py:100-110      (k2.rnnt_loss_smoothed)
 *   /root/reference/model/loss/pruned_rnnt_loss.py:39-48 (k2.rnnt_loss_pruned)
 *
 * k2 itself is an un-vendored dependency (requirements.txt:4 pins
 * k2==1.24.3.dev20240615+cuda11.6.torch1.13.1, Dockerfile.build:28-35 builds
 * tag v1.24.3); its source is not under /root/reference.  This file restates
 * the published algorithm of k2/python/csrc/torch/mutual_information_cpu.cu
 * (SURVEY.md Appendix A.2): a serial double loop per utterance, LogAdd with
 * the log(epsilon) cut-off, and the reverse recursion that yields the
 * occupation probabilities (px_grad, py_grad).
 *
 * "regular" RNN-T only: px is (B, S, T+1), py is (B, S+1, T), p is
 * (B, S+1, T+1), all contiguous; boundary is (B, 4) int64
 * [s_begin, t_begin, s_end, t_end].
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#define DEFINE_MI(SUFFIX, real_t, MIN_LOG_DIFF, EXP, LOG1P)                         \
  static inline real_t log_add_##SUFFIX(real_t x, real_t y) {                      \
    real_t diff;                                                                   \
    if (x < y) {                                                                   \
      diff = x - y;                                                                \
      x = y;                                                                       \
    } else {                                                                       \
      diff = y - x;                                                                \
    }                                                                              \
    /* diff <= 0 (or NaN when both are -inf, in which case x is returned) */       \
    if (diff >= (real_t)(MIN_LOG_DIFF)) return x + LOG1P(EXP(diff));               \
    return x;                                                                      \
  }                                                                                \
                                                                                   \
  /* forward: fills p (only inside the boundary rectangle) and ans[b] */           \
  void s2t_oracle_mi_forward_##SUFFIX(const real_t* px, const real_t* py,          \
                                      const int64_t* boundary, real_t* p,          \
                                      real_t* ans, int B, int S, int T,            \
                                      int b_begin, int b_end) {                    \
    const int64_t px_b = (int64_t)S * (T + 1), py_b = (int64_t)(S + 1) * T,        \
                  p_b = (int64_t)(S + 1) * (T + 1);                                \
    for (int b = b_begin; b < b_end && b < B; ++b) {                               \
      int s_begin = 0, t_begin = 0, s_end = S, t_end = T;                          \
      if (boundary) {                                                              \
        s_begin = (int)boundary[4 * b + 0];                                        \
        t_begin = (int)boundary[4 * b + 1];                                        \
        s_end = (int)boundary[4 * b + 2];                                          \
        t_end = (int)boundary[4 * b + 3];                                          \
      }                                                                            \
      const real_t* pxb = px + b * px_b;                                           \
      const real_t* pyb = py + b * py_b;                                           \
      real_t* pb = p + b * p_b;                                                    \
      const int T1 = T + 1;                                                        \
      pb[(int64_t)s_begin * T1 + t_begin] = (real_t)0;                             \
      for (int s = s_begin + 1; s <= s_end; ++s)                                   \
        pb[(int64_t)s * T1 + t_begin] = pb[(int64_t)(s - 1) * T1 + t_begin] +      \
                                        pxb[(int64_t)(s - 1) * T1 + t_begin];      \
      for (int t = t_begin + 1; t <= t_end; ++t)                                   \
        pb[(int64_t)s_begin * T1 + t] = pb[(int64_t)s_begin * T1 + t - 1] +        \
                                        pyb[(int64_t)s_begin * T + t - 1];         \
      for (int s = s_begin + 1; s <= s_end; ++s) {                                 \
        real_t p_s_t1 = pb[(int64_t)s * T1 + t_begin];                             \
        for (int t = t_begin + 1; t <= t_end; ++t) {                               \
          p_s_t1 = log_add_##SUFFIX(                                               \
              pb[(int64_t)(s - 1) * T1 + t] + pxb[(int64_t)(s - 1) * T1 + t],      \
              p_s_t1 + pyb[(int64_t)s * T + t - 1]);                               \
          pb[(int64_t)s * T1 + t] = p_s_t1;                                        \
        }                                                                          \
      }                                                                            \
      ans[b] = pb[(int64_t)s_end * T1 + t_end];                                    \
    }                                                                              \
  }                                                                                \
                                                                                   \
  /* backward: px_grad / py_grad must be zero-initialised by the caller;           \
     p_grad is a (B, S+1, T+1) scratch, also zero-initialised. */                  \
  void s2t_oracle_mi_backward_##SUFFIX(                                            \
      const real_t* px, const real_t* py, const int64_t* boundary,                 \
      const real_t* p, const real_t* ans_grad, real_t* p_grad, real_t* px_grad,    \
      real_t* py_grad, int B, int S, int T, int b_begin, int b_end) {              \
    const int64_t px_b = (int64_t)S * (T + 1), py_b = (int64_t)(S + 1) * T,        \
                  p_b = (int64_t)(S + 1) * (T + 1);                                \
    for (int b = b_begin; b < b_end && b < B; ++b) {                               \
      int s_begin = 0, t_begin = 0, s_end = S, t_end = T;                          \
      if (boundary) {                                                              \
        s_begin = (int)boundary[4 * b + 0];                                        \
        t_begin = (int)boundary[4 * b + 1];                                        \
        s_end = (int)boundary[4 * b + 2];                                          \
        t_end = (int)boundary[4 * b + 3];                                          \
      }                                                                            \
      const real_t* pxb = px + b * px_b;                                           \
      const real_t* pb = p + b * p_b;                                              \
      real_t* pgb = p_grad + b * p_b;                                              \
      real_t* pxg = px_grad + b * px_b;                                            \
      real_t* pyg = py_grad + b * py_b;                                            \
      const int T1 = T + 1;                                                        \
      (void)py;                                                                    \
      pgb[(int64_t)s_end * T1 + t_end] = ans_grad[b];                              \
      for (int s = s_end; s > s_begin; --s) {                                      \
        for (int t = t_end; t > t_begin; --t) {                                    \
          real_t term1 = pb[(int64_t)(s - 1) * T1 + t] +                           \
                         pxb[(int64_t)(s - 1) * T1 + t];                           \
          real_t total = pb[(int64_t)s * T1 + t];                                  \
          if (total - total != 0) total = 0;                                       \
          real_t term1_deriv = EXP(term1 - total);                                 \
          real_t term2_deriv = (real_t)1 - term1_deriv;                            \
          real_t grad = pgb[(int64_t)s * T1 + t];                                  \
          real_t term1_grad, term2_grad;                                           \
          if (term1_deriv - term1_deriv == 0) {                                    \
            term1_grad = term1_deriv * grad;                                       \
            term2_grad = term2_deriv * grad;                                       \
          } else {                                                                 \
            term1_grad = term2_grad = 0;                                           \
          }                                                                        \
          pxg[(int64_t)(s - 1) * T1 + t] = term1_grad;                             \
          pgb[(int64_t)(s - 1) * T1 + t] = term1_grad;                             \
          pyg[(int64_t)s * T + t - 1] = term2_grad;                                \
          pgb[(int64_t)s * T1 + t - 1] += term2_grad;                              \
        }                                                                          \
      }                                                                            \
      for (int t = t_end; t > t_begin; --t) {                                      \
        real_t g = pgb[(int64_t)s_begin * T1 + t];                                 \
        pgb[(int64_t)s_begin * T1 + t - 1] += g;                                   \
        pyg[(int64_t)s_begin * T + t - 1] = g;                                     \
      }                                                                            \
      for (int s = s_end; s > s_begin; --s) {                                      \
        real_t g = pgb[(int64_t)s * T1 + t_begin];                                 \
        pgb[(int64_t)(s - 1) * T1 + t_begin] += g;                                 \
        pxg[(int64_t)(s - 1) * T1 + t_begin] = g;                                  \
      }                                                                            \
    }                                                                              \
  }

/* log(FLT_EPSILON) = -15.9424, log(DBL_EPSILON) = -36.0437 */
DEFINE_MI(f32, float, -15.942385152878742, expf, log1pf)
DEFINE_MI(f64, double, -36.04365338911715, exp, log1p)

int s2t_oracle_abi_version(void) { return 1; }
