// Warp-synchronous lattice DPs: no CTA barrier inside the recursion.
//
// Same semantics as lattice.cu (k2's mutual_information forward/backward reached from
// /root/reference/model/joiner/joiner.py:100-110 and model/loss/pruned_rnnt_loss.py:39-48),
// restructured so that one anti-diagonal costs a couple of warp shuffles and one log-add-exp instead
// of a __syncthreads round trip (the barrier version spends ~1300 cycles per diagonal).
//
// (1) simple_lattice_*: the full (S+1) x (T+1) lattice in k2 layout.  One warp owns ALL symbol
//     positions of an utterance: lane l holds rows l, l+32, ... (RPL rows per lane), so the
//     neighbour value p(s-1, t) is a rotate-by-one shuffle and the RPL log-adds of a step are
//     independent (ILP).  alpha and beta run concurrently in separate CTAs.  Four producer warps
//     stream px/py from HBM with coalesced row reads and scatter them into a diagonal-major
//     shared-memory ring (odd row stride: conflict-free) that the recursion warp consumes through
//     mbarriers, one 32-diagonal block at a time.  alpha/beta go to HBM diagonal-major (coalesced)
//     and a third, fully parallel kernel forms the occupation probabilities.
// (2) band_lattice_kernel: the pruned band (B, T, R), R <= 32.  The whole band of an utterance fits
//     in shared memory; lane r owns band slot r and walks the frames, meeting its two predecessors
//     (slot r-1 of the same frame, slot r+delta of the previous frame) exactly one diagonal earlier,
//     i.e. in the neighbours' registers.  Warp 0 runs alpha while warp 1 runs beta; the CTA then
//     emits the occupation probabilities.
// Both keep running values relative to an fp64 offset re-based every 16 diagonals (see lattice.cu).
#include "lattice.cuh"
#include "tc_prims.cuh"

namespace s2t {
namespace {

using tc::mbar_arrive;
using tc::mbar_init;
using tc::mbar_wait;

constexpr int kRebaseShift = 4;  // offsets are constant over 16 consecutive diagonals
constexpr int kRingBlocks = 3;   // 32-diagonal blocks in the shared-memory ring

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMinLog2Diff = kMinLogDiff * kLog2e;  // k2's LogAdd cut-off, in the log2 domain

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// log2(2^x + 2^y): the recursions run in the log2 domain so that a log-add is MAX, SUB, EX2, ADD, LG2, ADD.
// Both -inf gives NaN in d, the compare fails and -inf (the max) comes back, as in k2's LogAdd.
__device__ __forceinline__ float log2_add(float x, float y) {
  const float mx = fmaxf(x, y), mn = fminf(x, y);
  const float d = mn - mx;
  return (d >= kMinLog2Diff) ? mx + __log2f(1.f + ex2_fast(d)) : mx;
}

// ------------------------------------------------------------------------------------------------
// (1) simple lattice
// ------------------------------------------------------------------------------------------------
struct SimpleArgs {
  const float* px;  // (B, S, T+1)
  const float* py;  // (B, S+1, T)
  const int64_t* boundary;
  int B, S, T;
  int rows_pad;       // 32 * RPL
  int diag_rows;      // diagonals allocated per utterance (multiple of 32)
  float* alpha_diag;  // (B, diag_rows, rows_pad), log2 domain, relative to aoff
  float* beta_diag;
  double* aoff;  // (B, n_off), log2 domain
  double* boff;
  int n_off;
  double* logp_d;  // (B), natural log
  float* logp;     // (B)
};

constexpr int kSimpleProducerWarps = 8;
constexpr int kSimpleThreads = 32 * (1 + kSimpleProducerWarps);

template <int RPL>
__global__ void __launch_bounds__(kSimpleThreads, 1) simple_lattice_kernel(SimpleArgs a) {
  constexpr int ROWS = 32 * RPL;
  constexpr int RS = ROWS + 1;  // odd stride: the producers' diagonal scatter is conflict-free
  constexpr int W = 32 * kRingBlocks;
  extern __shared__ float sm[];
  float* ringX = sm;
  float* ringY = sm + W * RS;
  uint64_t* full = reinterpret_cast<uint64_t*>(ringY + W * RS);  // 2 * W * RS floats: 8-byte aligned
  uint64_t* empty = full + kRingBlocks;

  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int Sb = a.S, Tb = a.T;
  if (a.boundary) {
    Sb = (int)a.boundary[4 * b + 2];
    Tb = (int)a.boundary[4 * b + 3];
  }
  Sb = min(max(Sb, 0), a.S);
  Tb = min(max(Tb, 0), a.T);
  const int nd = Sb + Tb;
  const int nblk = nd / 32 + 1;
  const float* px = a.px + (int64_t)b * a.S * (a.T + 1);
  const float* py = a.py + (int64_t)b * (a.S + 1) * a.T;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingBlocks; ++i) {
      mbar_init(&full[i], kSimpleProducerWarps);
      mbar_init(&empty[i], 1);
    }
    tc::fence_mbar_init();
  }
  __syncthreads();

  if (warp > 0) {
    // ---- producers: rows s = pw, pw + 8, ...; lane = diagonal inside the block.  They write EVERY
    // row of the ring (-inf outside the lattice) so that the recursion needs no masks.
    const int pw = warp - 1;
    for (int k = 0; k < nblk; ++k) {
      const int D = is_beta ? (nblk - 1 - k) : k;
      const int slot = k % kRingBlocks;
      mbar_wait(&empty[slot], ((k / kRingBlocks) & 1) ^ 1);
      const int d = 32 * D + lane;
      float* rx = ringX + (slot * 32 + lane) * RS;
      float* ry = ringY + (slot * 32 + lane) * RS;
      // all global loads of the block first (registers), then the shared-memory scatter: the compiler must
      // not be made to order a load behind a shared store it cannot prove independent
      constexpr int kRowsPerWarp = ROWS / kSimpleProducerWarps;
      float xs[kRowsPerWarp], ys[kRowsPerWarp];
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int s = pw + i * kSimpleProducerWarps;
        const int t = d - s;
        float xv = kNegInf, yv = kNegInf;
        if (s <= Sb) {
          if (!is_beta) {
            // alpha step into (s, t): X = px(s-1, t), Y = py(s, t-1)
            if (s >= 1 && t >= 0 && t <= Tb) xv = __ldg(px + (int64_t)(s - 1) * (a.T + 1) + t);
            if (t >= 1 && t <= Tb) yv = __ldg(py + (int64_t)s * a.T + t - 1);
          } else {
            // beta step out of (s, t): X = px(s, t), Y = py(s, t)
            if (s < Sb && t >= 0 && t <= Tb) xv = __ldg(px + (int64_t)s * (a.T + 1) + t);
            if (t >= 0 && t < Tb) yv = __ldg(py + (int64_t)s * a.T + t);
          }
        }
        xs[i] = xv;
        ys[i] = yv;
      }
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int s = pw + i * kSimpleProducerWarps;
        rx[s] = kLog2e * xs[i];
        ry[s] = kLog2e * ys[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    }
    return;
  }

  // ---- recursion warp: lane l holds rows l, l + 32, ... ----
  float v[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) v[j] = kNegInf;
  double off = 0.0;
  float* out = (is_beta ? a.beta_diag : a.alpha_diag) + (int64_t)b * a.diag_rows * ROWS;
  double* offs = (is_beta ? a.boff : a.aoff) + (int64_t)b * a.n_off;
  const int src_lane = is_beta ? ((lane + 1) & 31) : ((lane + 31) & 31);
  const int fin_j = Sb >> 5, fin_lane = Sb & 31;
  float fin = kNegInf;
  double fin_off = 0.0;

  for (int k = 0; k < nblk; ++k) {
    const int D = is_beta ? (nblk - 1 - k) : k;
    const int slot = k % kRingBlocks;
    mbar_wait(&full[slot], (k / kRingBlocks) & 1);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {  // two 16-diagonal periods per block
      const int dh = is_beta ? (32 * D + 31 - 16 * h) : (32 * D + 16 * h);  // first diagonal of the period
      if (lane == 0) offs[dh >> kRebaseShift] = off;
#pragma unroll 4
      for (int ii = 0; ii < 16; ++ii) {
        const int d = is_beta ? dh - ii : dh + ii;
        const float* rx = ringX + (slot * 32 + (d & 31)) * RS;
        const float* ry = ringY + (slot * 32 + (d & 31)) * RS;
        float rot[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) rot[j] = __shfl_sync(0xffffffffu, v[j], src_lane);
        float nv[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          float nb;  // neighbour row s-1 (alpha) / s+1 (beta) on the previous diagonal
          if (!is_beta) nb = (lane > 0) ? rot[j] : (j > 0 ? rot[j - 1] : kNegInf);
          else nb = (lane < 31) ? rot[j] : (j < RPL - 1 ? rot[j + 1] : kNegInf);
          nv[j] = log2_add(nb + rx[j * 32 + lane], v[j] + ry[j * 32 + lane]);
        }
        // the single source cell: alpha(0, 0) = 0 on diagonal 0, beta(S_b, T_b) = 0 on diagonal nd
        if (!is_beta) {
          if (d == 0 && lane == 0) nv[0] = 0.f;
        } else if (d == nd && lane == fin_lane) {
#pragma unroll
          for (int j = 0; j < RPL; ++j)
            if (j == fin_j) nv[j] = 0.f;
        }
        float* o = out + (int64_t)d * ROWS + lane;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          v[j] = nv[j];
          o[j * 32] = nv[j];
        }
        if (!is_beta && d == nd) {  // log P(y|x) = alpha(S_b, T_b)
          float mine = kNegInf;
#pragma unroll
          for (int j = 0; j < RPL; ++j)
            if (j == fin_j) mine = v[j];
          fin = __shfl_sync(0xffffffffu, mine, fin_lane);
          fin_off = off;
        }
      }
      float m = v[0];
#pragma unroll
      for (int j = 1; j < RPL; ++j) m = fmaxf(m, v[j]);
      m = warp_max(m);
      if (m - m == 0.f) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) v[j] -= m;
        off += (double)m;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
  if (!is_beta && lane == 0) {
    const double lp = ((double)fin + fin_off) * (double)kLn2;
    a.logp_d[b] = lp;
    a.logp[b] = (float)lp;
  }
}

// occupation probabilities in k2 layout from diagonal-major alpha / beta; writes EVERY element of
// px_grad (B,S,T+1) and py_grad (B,S+1,T) (zeros outside the live region: no memset needed)
__global__ void __launch_bounds__(256) simple_occupation_kernel(SimpleArgs a, float* __restrict__ occ_px,
                                                                float* __restrict__ occ_py) {
  __shared__ float sA[63][33];
  __shared__ float sB[63][34];
  const int b = blockIdx.z;
  const int s0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int ROWS = a.rows_pad;
  int Sb = a.S, Tb = a.T;
  if (a.boundary) {
    Sb = (int)a.boundary[4 * b + 2];
    Tb = (int)a.boundary[4 * b + 3];
  }
  Sb = min(max(Sb, 0), a.S);
  Tb = min(max(Tb, 0), a.T);
  const int nd = Sb + Tb;
  const float* ad = a.alpha_diag + (int64_t)b * a.diag_rows * ROWS;
  const float* bd = a.beta_diag + (int64_t)b * a.diag_rows * ROWS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int d0 = s0 + t0;
  const bool tile_live = (s0 <= Sb) && (t0 <= Tb);
  if (tile_live) {
    for (int r = ty; r < 63; r += 8) {
      const int d = d0 + r;
      const int s = s0 + tx;
      sA[r][tx] = (d <= nd && s < ROWS) ? ad[(int64_t)d * ROWS + s] : kNegInf;
      sB[r][tx] = (d + 1 <= nd && s < ROWS) ? bd[(int64_t)(d + 1) * ROWS + s] : kNegInf;
      if (tx == 0) sB[r][32] = (d + 1 <= nd && s0 + 32 < ROWS) ? bd[(int64_t)(d + 1) * ROWS + s0 + 32] : kNegInf;
    }
  }
  __syncthreads();
  const double lp = a.logp_d[b];
  const bool lp_ok = (lp - lp == 0.0);
  const double* aoff = a.aoff + (int64_t)b * a.n_off;
  const double* boff = a.boff + (int64_t)b * a.n_off;
  for (int rs = ty; rs < 32; rs += 8) {
    const int s = s0 + rs, t = t0 + tx;
    if (s > a.S || t > a.T) continue;
    float ox = 0.f, oy = 0.f;
    if (tile_live && lp_ok && s <= Sb && t <= Tb) {
      const int d = s + t;
      // alpha / beta and their offsets are in the log2 domain
      const float cst = (float)(aoff[d >> kRebaseShift] + boff[(d + 1 <= nd ? d + 1 : nd) >> kRebaseShift] -
                                lp * (double)kLog2e);
      const float av = sA[rs + tx][rs];
      if (t < Tb) {
        const float yv = kLog2e * __ldg(a.py + ((int64_t)b * (a.S + 1) + s) * a.T + t);
        oy = exp2f(av + yv + sB[rs + tx][rs] + cst);
      }
      if (s < Sb) {
        const float xv = kLog2e * __ldg(a.px + ((int64_t)b * a.S + s) * (a.T + 1) + t);
        ox = exp2f(av + xv + sB[rs + tx][rs + 1] + cst);
      }
    }
    if (s < a.S) occ_px[((int64_t)b * a.S + s) * (a.T + 1) + t] = ox;
    if (t < a.T) occ_py[((int64_t)b * (a.S + 1) + s) * a.T + t] = oy;
  }
}

// ------------------------------------------------------------------------------------------------
// (2) pruned band
// ------------------------------------------------------------------------------------------------
struct BandArgs {
  const float* px;  // (B, T, R)
  const float* py;
  const int64_t* ranges;  // (B, T, R)
  const int64_t* boundary;
  int B, S, T, R;
  float* logp;
  float* occ_px;  // (B, T, R), fully written
  float* occ_py;
};

__global__ void __launch_bounds__(128, 1) band_lattice_kernel(BandArgs a) {
  extern __shared__ float sm[];
  const int T = a.T, R = a.R;
  float* pxs = sm;            // log2 domain
  float* pys = pxs + T * R;
  float* als = pys + T * R;   // alpha / beta relative to the offset of their diagonal
  float* bes = als + T * R;
  float* yis = bes + T * R;   // alpha's incoming blank move: py(t-1, r + sb[t] - sb[t-1]) or -inf
  int* sbs = reinterpret_cast<int*>(yis + T * R);  // sb[t], T + 2 entries (padding frames repeat the last)
  double* offs = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sbs + T + 2) + 7) & ~(uintptr_t)7);
  const int n_off = ((a.S + T + R) >> kRebaseShift) + 2;
  double* aoff = offs;
  double* boff = offs + n_off;
  __shared__ double s_logp2;  // log2 P(y|x)

  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int Sb = a.S, Tb = T;
  if (a.boundary) {
    Sb = (int)a.boundary[4 * b + 2];
    Tb = (int)a.boundary[4 * b + 3];
  }
  Sb = min(max(Sb, 0), a.S);
  Tb = min(max(Tb, 0), T);
  const float* gpx = a.px + (int64_t)b * T * R;
  const float* gpy = a.py + (int64_t)b * T * R;
  const int64_t* grg = a.ranges ? a.ranges + (int64_t)b * T * R : nullptr;

  for (int i = threadIdx.x; i < Tb * R; i += blockDim.x) {
    pxs[i] = kLog2e * __ldg(gpx + i);
    pys[i] = kLog2e * __ldg(gpy + i);
  }
  for (int t = threadIdx.x; t <= Tb + 1 && t < T + 2; t += blockDim.x) {
    const int tt = min(t, max(Tb - 1, 0));
    sbs[t] = grg ? (int)grg[(int64_t)tt * R] : 0;
  }
  if (threadIdx.x == 0) s_logp2 = -INFINITY;
  __syncthreads();
  for (int i = threadIdx.x; i < Tb * R; i += blockDim.x) {
    const int t = i / R, r = i - t * R;
    float v = kNegInf;
    if (t > 0) {
      const int r2 = r + sbs[t] - sbs[t - 1];
      if (r2 < R) v = pys[(t - 1) * R + r2];
    }
    yis[i] = v;
  }
  __syncthreads();

  constexpr int kDone = 1 << 29;
  const int last_d = (Tb > 0) ? (Tb - 1 + sbs[Tb - 1] + R - 1) : -1;  // last diagonal that holds a band cell
  const int n_periods = (last_d >> kRebaseShift) + 1;
  if (warp == 0 && Tb > 0) {
    // ---- alpha: lane r owns slot r and walks the frames upwards.  Per frame it pre-loads the two
    // incoming log-probs (already -inf when the move or the cell does not exist), so a step is two
    // shuffles and one log2-add.
    // Everything a frame needs is fetched one frame ahead (independent shared-memory loads), so the only
    // dependent chain of a step is shuffle -> log2-add.
    int t = 0, sb_cur = sbs[0];
    int f = (lane < R) ? sb_cur + lane : kDone;  // diagonal of this lane's next cell
    int delta = 0;
    const bool act = lane < R;
    float xin = (lane > 0 && act) ? pxs[lane - 1] : kNegInf;  // px(t, r-1)
    float yin = kNegInf;                                       // py(t-1, r+delta): none for t = 0
    int tn = min(1, Tb - 1);
    int sb_n = sbs[tn];
    float xin_n = (lane > 0 && act) ? pxs[tn * R + lane - 1] : kNegInf;
    float yin_n = act ? yis[tn * R + lane] : kNegInf;
    float* ap = als + lane;
    float last = kNegInf;
    double off = 0.0;
    for (int p = 0; p < n_periods; ++p) {
      if (lane == 0) aoff[p] = off;
#pragma unroll 4
      for (int ii = 0; ii < 16; ++ii) {
        const int d = 16 * p + ii;
        const float up_src = __shfl_up_sync(0xffffffffu, last, 1);
        const float left_src = __shfl_sync(0xffffffffu, last, (lane + delta) & 31);
        float val = kNegInf;
        if (f == d) {
          val = log2_add(up_src + xin, left_src + yin);
          if (d == 0) val = (lane == 0) ? 0.f : kNegInf;  // alpha(0, 0) = 0
          if (sb_cur + lane > Sb) val = kNegInf;
          *ap = val;
          ap += R;
          ++t;
          if (t < Tb) {
            delta = sb_n - sb_cur;
            sb_cur = sb_n;
            f = t + sb_cur + lane;
            xin = xin_n;
            yin = yin_n;
            tn = min(t + 1, Tb - 1);
            sb_n = sbs[tn];
            xin_n = (lane > 0) ? pxs[tn * R + lane - 1] : kNegInf;
            yin_n = yis[tn * R + lane];
          } else {
            f = kDone;
          }
        }
        last = val;
      }
      const float m = warp_max(last);  // cells older than this diagonal already sit in als[]
      if (m - m == 0.f) {
        last -= m;
        off += (double)m;
      }
    }
  } else if (warp == 1 && Tb > 0) {
    // ---- beta: lane r owns slot r and walks the frames downwards ----
    int t = Tb - 1, sb_cur = sbs[Tb - 1];
    int f = (lane < R) ? t + sb_cur + lane : -kDone;
    int delta = 0;  // sb[t+1] - sb[t]
    const bool act = lane < R;
    const int s0 = sb_cur + lane;
    float xout = (lane + 1 < R && s0 < Sb) ? pxs[t * R + lane] : kNegInf;  // px(t, r) -> slot r+1
    float yout = (act && s0 == Sb) ? pys[t * R + lane] : kNegInf;         // last frame: beta(s, T_b) = [s == S_b]
    bool term = true;  // the blank move of the last frame ends the lattice
    int tp = max(t - 1, 0);  // frame fetched ahead
    int sb_p = sbs[tp];
    float xraw_p = act ? pxs[tp * R + lane] : kNegInf;
    float yraw_p = act ? pys[tp * R + lane] : kNegInf;
    float* bp = bes + t * R + lane;
    float last = kNegInf;
    double off = 0.0;
    for (int p = n_periods - 1; p >= 0; --p) {
      if (lane == 0) boff[p] = off;
#pragma unroll 4
      for (int ii = 15; ii >= 0; --ii) {
        const int d = 16 * p + ii;
        const float down_src = __shfl_down_sync(0xffffffffu, last, 1);
        const float right_src = __shfl_sync(0xffffffffu, last, (lane - delta) & 31);
        float val = kNegInf;
        if (f == d) {
          val = log2_add(down_src + xout, term ? yout : right_src + yout);
          if (sb_cur + lane > Sb) val = kNegInf;
          *bp = val;
          bp -= R;
          --t;
          term = false;
          if (t >= 0) {
            delta = sb_cur - sb_p;
            sb_cur = sb_p;
            f = t + sb_cur + lane;
            xout = (lane + 1 < R && sb_cur + lane < Sb) ? xraw_p : kNegInf;
            yout = (lane - delta >= 0) ? yraw_p : kNegInf;
            tp = max(t - 1, 0);
            sb_p = sbs[tp];
            xraw_p = pxs[tp * R + lane];
            yraw_p = pys[tp * R + lane];
          } else {
            f = -kDone;
          }
        }
        last = val;
      }
      const float m = warp_max(last);
      if (m - m == 0.f) {
        last -= m;
        off += (double)m;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double lp2 = -INFINITY;
    if (Tb > 0) {
      // log P = alpha(T_b-1, r*) + py(T_b-1, r*), r* = S_b - sb[T_b-1]
      const int rs = Sb - sbs[Tb - 1];
      if (rs >= 0 && rs < R) {
        const int dd = Tb - 1 + sbs[Tb - 1] + rs;
        lp2 = (double)als[(Tb - 1) * R + rs] + (double)pys[(Tb - 1) * R + rs] + aoff[dd >> kRebaseShift];
      }
    } else if (Sb == 0) {
      lp2 = 0.0;  // no frames: only the empty transcript is possible
    }
    s_logp2 = lp2;
    a.logp[b] = (float)(lp2 * (double)kLn2);
  }
  __syncthreads();
  const double lp2 = s_logp2;
  const bool lp_ok = (lp2 - lp2 == 0.0) && Tb > 0;
  float* ox = a.occ_px + (int64_t)b * T * R;
  float* oy = a.occ_py + (int64_t)b * T * R;
  for (int i = threadIdx.x; i < T * R; i += blockDim.x) {
    const int t = i / R, r = i - t * R;
    float vx = 0.f, vy = 0.f;
    if (lp_ok && t < Tb) {
      const int s = sbs[t] + r;
      if (s <= Sb) {
        const int d = t + sbs[t] + r;
        const float av = als[i];
        const double ao = aoff[d >> kRebaseShift];
        if (t == Tb - 1) {
          if (s == Sb) vy = exp2f((float)((double)av + (double)pys[i] + ao - lp2));
        } else {
          const int r2 = s - sbs[t + 1];  // blank: to (s, t+1)
          if (r2 >= 0 && r2 < R) {
            const float cst = (float)(ao + boff[(d + 1) >> kRebaseShift] - lp2);
            vy = exp2f(av + pys[i] + bes[(t + 1) * R + r2] + cst);
          }
        }
        if (r + 1 < R && s < Sb) {  // symbol: to (s+1, t)
          const float cst = (float)(ao + boff[(d + 1) >> kRebaseShift] - lp2);
          vx = exp2f(av + pxs[i] + bes[t * R + r + 1] + cst);
        }
      }
    }
    ox[i] = vx;
    oy[i] = vy;
  }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
static int simple_rpl(int S) {
  const int rows = S + 1;
  if (rows <= 32) return 1;
  if (rows <= 64) return 2;
  if (rows <= 128) return 4;
  if (rows <= 256) return 8;
  return 0;
}

bool simple_lattice_fast_ok(int S) { return simple_rpl(S) > 0; }

size_t simple_lattice_fast_workspace_bytes(int B, int S, int T) {
  const int rpl = simple_rpl(S);
  if (!rpl) return 0;
  const size_t diag_rows = (size_t)((S + T) / 32 + 1) * 32;
  const size_t diag = (size_t)B * diag_rows * 32 * rpl * sizeof(float);
  const size_t n_off = diag_rows / 16 + 2;
  return 2 * diag + (2 * (size_t)B * n_off + B) * sizeof(double) + 64;
}

int launch_simple_lattice_fast(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                               void* ws, float* logp, float* occ_px, float* occ_py, cudaStream_t stream) {
  const int rpl = simple_rpl(S);
  S2T_REQUIRE(rpl > 0, "simple_lattice_fast: S+1 = %d > 256", S + 1);
  if (B == 0) return 0;
  SimpleArgs a{};
  a.px = px; a.py = py; a.boundary = boundary;
  a.B = B; a.S = S; a.T = T;
  a.rows_pad = 32 * rpl;
  a.diag_rows = ((S + T) / 32 + 1) * 32;
  const size_t diag = (size_t)B * a.diag_rows * a.rows_pad;
  a.n_off = a.diag_rows / 16 + 2;
  a.alpha_diag = (float*)ws;
  a.beta_diag = a.alpha_diag + diag;
  a.aoff = (double*)(a.beta_diag + diag + ((2 * diag) & 1));
  a.boff = a.aoff + (size_t)B * a.n_off;
  a.logp_d = a.boff + (size_t)B * a.n_off;
  a.logp = logp;
  const size_t smem = (size_t)2 * 32 * kRingBlocks * (a.rows_pad + 1) * sizeof(float) + 8 + 2 * kRingBlocks * 8 + 16;
  {
    ProfScope prof("simple_lattice_kernel", stream);
    dim3 grid(B, occ_px ? 2 : 1);
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kSimpleThreads, smem, stream>>>(a);
    };
    switch (rpl) {
      case 1: launch(simple_lattice_kernel<1>); break;
      case 2: launch(simple_lattice_kernel<2>); break;
      case 4: launch(simple_lattice_kernel<4>); break;
      default: launch(simple_lattice_kernel<8>); break;
    }
  }
  if (int rc = check_launch("simple_lattice_kernel")) return rc;
  if (occ_px) {
    ProfScope prof("simple_occupation_kernel", stream);
    dim3 grid((T + 1 + 31) / 32, (S + 1 + 31) / 32, B);
    simple_occupation_kernel<<<grid, 256, 0, stream>>>(a, occ_px, occ_py);
    return check_launch("simple_occupation_kernel");
  }
  return 0;
}

static size_t band_smem_bytes(int S, int T, int R) {
  const size_t n_off = ((S + T + R) >> kRebaseShift) + 2;
  return (size_t)5 * T * R * sizeof(float) + (size_t)(T + 2) * sizeof(int) + 8 + 2 * n_off * sizeof(double) + 16;
}

bool band_lattice_fast_ok(int S, int T, int R) { return R <= 32 && band_smem_bytes(S, T, R) <= 200 * 1024; }

int launch_band_lattice_fast(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary,
                             int B, int S, int T, int R, float* logp, float* occ_px, float* occ_py,
                             cudaStream_t stream) {
  if (B == 0) return 0;
  BandArgs a{px, py, ranges, boundary, B, S, T, R, logp, occ_px, occ_py};
  const size_t smem = band_smem_bytes(S, T, R);
  cudaFuncSetAttribute(band_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ProfScope prof("band_lattice_kernel", stream);
  band_lattice_kernel<<<B, 128, smem, stream>>>(a);
  return check_launch("band_lattice_kernel");
}

}  // namespace s2t
