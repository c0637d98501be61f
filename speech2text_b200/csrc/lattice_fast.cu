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
  int groups;         // row groups of 32 symbol positions (= recursion warps)
  int rows_pad;       // 32 * groups
  int diag_rows;      // diagonals allocated per utterance (multiple of 32)
  float* alpha_diag;  // (B, diag_rows, rows_pad), log2 domain, relative to aoff of the row's group
  float* beta_diag;
  double* aoff;  // (B, groups, n_off), log2 domain
  double* boff;
  int n_off;
  double* logp_d;  // (B), natural log
  float* logp;     // (B)
};

constexpr int kSimpleProducerWarps = 8;

// Systolic layout: recursion warp g owns symbol positions 32 g .. 32 g + 31 (one per lane), so a diagonal
// step is ONE shuffle + ONE log2-add per warp.  The value a group needs from its neighbour group (row
// 32 g - 1 for alpha, 32 g + 32 for beta, one diagonal earlier) travels through a small shared-memory
// array; group g runs one 32-diagonal block behind the neighbour it depends on (mbarrier per block), so
// the G warps work on G different blocks at once, like a pipeline.  Every group keeps its own fp64 offset
// (re-based every 16 diagonals by the group's maximum); boundary values are converted between the offsets
// of the two groups.  Producer warps (8 / G per group) stream px / py from HBM into a per-group
// diagonal-major ring.
struct SimpleSmem {
  float* ringX;   // [G][W * RS]
  float* ringY;   // [G][W * RS]
  float* bnd;     // [G + 1][bnd_stride]: boundary-row value per diagonal at index d + 1; array G stays -inf
  double* offs;   // [G][n_off]
  uint64_t* full;   // [G][kRing]
  uint64_t* empty;  // [G][kRing]
  uint64_t* bfull;  // [G][nblk_max]
  int bnd_stride;
};

template <int G, bool kBeta>
__device__ __forceinline__ void simple_recursion(const SimpleArgs& a, const SimpleSmem& sh, int b, int g, int lane,
                                                 int Sb, int nd, int nblk) {
  constexpr int ROWS = 32 * G;
  constexpr int RS = 33;
  constexpr int kRing = G <= 4 ? 3 : 2;
  constexpr int W = 32 * kRing;
  constexpr int kYOff = G * W * RS;  // ringY - ringX
  const int nblk_max = a.diag_rows / 32;
  const int dep = kBeta ? g + 1 : g - 1;  // the group whose boundary row feeds this one
  const bool has_dep = dep >= 0 && dep < G;
  const int edge_lane = kBeta ? 31 : 0;   // lane that takes its neighbour from the other group
  const int send_lane = kBeta ? 0 : 31;   // lane whose value the next group needs
  const float* gx = sh.ringX + g * W * RS + lane;
  float* my_bnd = sh.bnd + g * sh.bnd_stride + 1;
  const float* dep_bnd = sh.bnd + (has_dep ? dep : G) * sh.bnd_stride + 1;
  double* my_offs = sh.offs + g * a.n_off;
  const double* dep_offs = sh.offs + (has_dep ? dep : 0) * a.n_off;
  double* g_offs = (kBeta ? a.boff : a.aoff) + ((int64_t)b * G + g) * a.n_off;
  float* out0 = (kBeta ? a.beta_diag : a.alpha_diag) + (int64_t)b * a.diag_rows * ROWS + 32 * g + lane;
  const int src_lane = kBeta ? ((lane + 1) & 31) : ((lane + 31) & 31);
  const bool fin_mine = (Sb >> 5) == g;
  const int fin_lane = Sb & 31;
  constexpr int kStep = kBeta ? -1 : 1;
  float v = kNegInf;
  double off = 0.0;

  for (int k = 0; k < nblk; ++k) {
    const int D = kBeta ? (nblk - 1 - k) : k;
    const int slot = k % kRing;
    mbar_wait(&sh.full[g * kRing + slot], (k / kRing) & 1);
    if (has_dep) mbar_wait(&sh.bfull[dep * nblk_max + k], 0);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {  // two 16-diagonal periods per block
      const int dh = kBeta ? (32 * D + 31 - 16 * h) : (32 * D + 16 * h);  // first diagonal of the period
      const int p = dh >> kRebaseShift;
      if (lane == 0) {
        my_offs[p] = off;
        g_offs[p] = off;
      }
      // Boundary values of this period, one per lane (lane i serves step i), converted from the neighbour
      // group's offset of the period they were computed in to this group's current offset.
      float bvals = kNegInf;
      if (has_dep && lane < 16) {
        const int pp = kBeta ? p + 1 : p - 1;
        const bool pp_ok = kBeta ? (16 * pp < 32 * nblk) : (pp >= 0);
        const double o = (lane == 0) ? (pp_ok ? dep_offs[pp] : off) : dep_offs[p];
        const int dp = kBeta ? dh - lane + 1 : dh + lane - 1;
        bvals = dep_bnd[dp] + (float)(o - off);
      }
      const float* rx = gx + (slot * 32 + (dh & 31)) * RS;
      float* out = out0 + (int64_t)dh * ROWS;
      float* bp = my_bnd + dh;
      const bool special = kBeta ? (nd <= dh && nd > dh - 16) : (dh == 0 || (nd >= dh && nd < dh + 16));
      if (!special) {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) {
          const float rot = __shfl_sync(0xffffffffu, v, src_lane);
          const float bv = __shfl_sync(0xffffffffu, bvals, ii);
          const float nb = (lane == edge_lane) ? bv : rot;
          v = log2_add(nb + rx[kStep * ii * RS], v + rx[kStep * ii * RS + kYOff]);
          out[kStep * ii * ROWS] = v;
          if (lane == send_lane) bp[kStep * ii] = v;
        }
      } else {
#pragma unroll 4
        for (int ii = 0; ii < 16; ++ii) {
          const int d = dh + kStep * ii;
          const float rot = __shfl_sync(0xffffffffu, v, src_lane);
          const float bv = __shfl_sync(0xffffffffu, bvals, ii);
          const float nb = (lane == edge_lane) ? bv : rot;
          float nv = log2_add(nb + rx[kStep * ii * RS], v + rx[kStep * ii * RS + kYOff]);
          // the single source cell: alpha(0, 0) = 0 on diagonal 0, beta(S_b, T_b) = 0 on diagonal nd
          if (!kBeta) {
            if (d == 0 && g == 0 && lane == 0) nv = 0.f;
          } else if (d == nd && fin_mine && lane == fin_lane) {
            nv = 0.f;
          }
          v = nv;
          out[kStep * ii * ROWS] = v;
          if (lane == send_lane) bp[kStep * ii] = v;
          if (!kBeta && d == nd && fin_mine) {  // log P(y|x) = alpha(S_b, T_b)
            const float fin = __shfl_sync(0xffffffffu, v, fin_lane);
            if (lane == 0) {
              const double lp = ((double)fin + off) * (double)kLn2;
              a.logp_d[b] = lp;
              a.logp[b] = (float)lp;
            }
          }
        }
      }
      const float m = warp_max(v);
      if (m - m == 0.f) {
        v -= m;
        off += (double)m;
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&sh.empty[g * kRing + slot]);
      mbar_arrive(&sh.bfull[g * nblk_max + k]);
    }
  }
}

template <int G>
__global__ void __launch_bounds__(32 * (G + kSimpleProducerWarps), 1) simple_lattice_kernel(SimpleArgs a) {
  constexpr int RS = 33;                 // odd row stride: the producers' diagonal scatter is conflict-free
  constexpr int kRing = G <= 4 ? 3 : 2;  // 32-diagonal blocks per group ring
  constexpr int W = 32 * kRing;
  constexpr int kProdPerGroup = kSimpleProducerWarps / G;
  constexpr int kRowsPerProd = 32 / kProdPerGroup;
  extern __shared__ float sm[];
  const int nblk_max = a.diag_rows / 32;
  SimpleSmem sh;
  sh.bnd_stride = a.diag_rows + 2;
  sh.ringX = sm;
  sh.ringY = sh.ringX + G * W * RS;
  sh.bnd = sh.ringY + G * W * RS;
  sh.offs = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sh.bnd + (G + 1) * sh.bnd_stride) + 7) & ~(uintptr_t)7);
  sh.full = reinterpret_cast<uint64_t*>(sh.offs + G * a.n_off);
  sh.empty = sh.full + G * kRing;
  sh.bfull = sh.empty + G * kRing;

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

  for (int i = threadIdx.x; i < G * kRing; i += blockDim.x) {
    mbar_init(&sh.full[i], kProdPerGroup);
    mbar_init(&sh.empty[i], 1);
  }
  for (int i = threadIdx.x; i < G * nblk_max; i += blockDim.x) mbar_init(&sh.bfull[i], 1);
  for (int i = threadIdx.x; i < (G + 1) * sh.bnd_stride; i += blockDim.x) sh.bnd[i] = kNegInf;
  for (int i = threadIdx.x; i < G * a.n_off; i += blockDim.x) sh.offs[i] = 0.0;
  tc::fence_mbar_init();
  __syncthreads();

  if (warp >= G) {
    // ---- producers: lane = diagonal inside the block.  They write EVERY row of the ring (-inf outside
    // the lattice) so that the recursion needs no masks.
    const int pw = warp - G;
    const int g = pw % G, sub = pw / G;
    float* gx = sh.ringX + g * W * RS;
    float* gy = sh.ringY + g * W * RS;
    for (int k = 0; k < nblk; ++k) {
      const int D = is_beta ? (nblk - 1 - k) : k;
      const int slot = k % kRing;
      mbar_wait(&sh.empty[g * kRing + slot], ((k / kRing) & 1) ^ 1);
      const int d = 32 * D + lane;
      float* rx = gx + (slot * 32 + lane) * RS;
      float* ry = gy + (slot * 32 + lane) * RS;
      // all global loads of the block first (registers), then the shared-memory scatter
      float xs[kRowsPerProd], ys[kRowsPerProd];
#pragma unroll
      for (int i = 0; i < kRowsPerProd; ++i) {
        const int s = 32 * g + sub + i * kProdPerGroup;
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
      for (int i = 0; i < kRowsPerProd; ++i) {
        const int r = sub + i * kProdPerGroup;
        rx[r] = kLog2e * xs[i];
        ry[r] = kLog2e * ys[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.full[g * kRing + slot]);
    }
    return;
  }
  // ---- recursion warp g: lane l holds symbol position 32 g + l ----
  if (is_beta) simple_recursion<G, true>(a, sh, b, warp, lane, Sb, nd, nblk);
  else simple_recursion<G, false>(a, sh, b, warp, lane, Sb, nd, nblk);
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
  const double* aoff = a.aoff + (int64_t)b * a.groups * a.n_off;  // [group][period]
  const double* boff = a.boff + (int64_t)b * a.groups * a.n_off;
  for (int rs = ty; rs < 32; rs += 8) {
    const int s = s0 + rs, t = t0 + tx;
    if (s > a.S || t > a.T) continue;
    float ox = 0.f, oy = 0.f;
    if (tile_live && lp_ok && s <= Sb && t <= Tb) {
      const int d = s + t;
      // alpha / beta and the offsets of their row groups are in the log2 domain
      const int pb = (d + 1 <= nd ? d + 1 : nd) >> kRebaseShift;
      const double ca = aoff[(s >> 5) * a.n_off + (d >> kRebaseShift)] - lp * (double)kLog2e;
      const float av = sA[rs + tx][rs];
      if (t < Tb) {
        const float yv = kLog2e * __ldg(a.py + ((int64_t)b * (a.S + 1) + s) * a.T + t);
        oy = exp2f(av + yv + sB[rs + tx][rs] + (float)(ca + boff[(s >> 5) * a.n_off + pb]));
      }
      if (s < Sb) {
        const float xv = kLog2e * __ldg(a.px + ((int64_t)b * a.S + s) * (a.T + 1) + t);
        ox = exp2f(av + xv + sB[rs + tx][rs + 1] + (float)(ca + boff[((s + 1) >> 5) * a.n_off + pb]));
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

// One step record per band cell and direction, built in parallel before the recursions start:
//   x, y   incoming (alpha) / outgoing (beta) log2-probabilities, -inf where the move or the cell does not exist
//   f      diagonal t + sb[t] + r of the cell (kBandDone / -kBandDone after the last cell of the lane)
//   w      bits 0..7: lane distance delta to the blank-move neighbour; bit 8: the blank-move term is a
//          constant 0 (alpha(0,0) = 0 and the lattice exit of the last frame) instead of a neighbour value
struct __align__(16) BandStep {
  float x, y;
  int f, w;
};
constexpr int kBandDone = 1 << 29;

__global__ void __launch_bounds__(128, 1) band_lattice_kernel(BandArgs a) {
  extern __shared__ float sm[];
  const int T = a.T, R = a.R;
  float* pxs = sm;            // log2 domain
  float* pys = pxs + T * R;
  float* als = pys + T * R;   // alpha / beta relative to the offset of their diagonal
  float* bes = als + T * R;
  BandStep* arec = reinterpret_cast<BandStep*>((reinterpret_cast<uintptr_t>(bes + T * R) + 15) & ~(uintptr_t)15);  // (T + 1) * R
  BandStep* brec = arec + (T + 1) * R + 3 * R;  // frame t at brec + t * R; frames -3..-1 are padding / terminator
  int* sbs = reinterpret_cast<int*>(brec + (T + 1) * R);  // sb[t], T + 2 entries (padding frames repeat the last)
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
  for (int i = threadIdx.x; i < (Tb + 1) * R; i += blockDim.x) {
    const int t = i / R, r = i - t * R;
    BandStep ra{kNegInf, kNegInf, kBandDone, 0}, rb{kNegInf, kNegInf, -kBandDone, 0};
    if (t < Tb) {
      const int sb = sbs[t], s = sb + r;
      const bool in = s <= Sb;
      // alpha into (t, r): symbol move from (t, r-1), blank move from (t-1, r + sb[t] - sb[t-1])
      const int da = t > 0 ? sb - sbs[t - 1] : 0;
      ra.f = t + sb + r;
      ra.w = da & 255;
      if (in) {
        if (r > 0) ra.x = pxs[i - 1];
        if (t > 0 && r + da < R) ra.y = pys[(t - 1) * R + r + da];
        if (t == 0 && s == 0) {  // alpha(0, 0) = 0
          ra.x = kNegInf;
          ra.y = 0.f;
          ra.w |= 256;
        }
      }
      // beta out of (t, r): symbol move to (t, r+1), blank move to (t+1, r - (sb[t+1] - sb[t])) or out of the lattice
      const int db = t + 1 < Tb ? sbs[t + 1] - sb : 0;
      rb.f = t + sb + r;
      rb.w = db & 255;
      if (in) {
        if (r + 1 < R && s < Sb) rb.x = pxs[i];
        if (t + 1 < Tb) {
          if (r - db >= 0) rb.y = pys[i];
        } else {
          rb.w |= 256;  // last frame: beta(s, T_b) = [s == S_b]
          if (s == Sb) rb.y = pys[i];
        }
      }
    }
    arec[i] = ra;
    brec[i] = rb;
  }
  for (int i = threadIdx.x; i < 3 * R; i += blockDim.x) brec[i - 3 * R] = BandStep{kNegInf, kNegInf, -kBandDone, 0};
  __syncthreads();

  const int last_d = (Tb > 0) ? (Tb - 1 + sbs[Tb - 1] + R - 1) : -1;  // last diagonal that holds a band cell
  const int n_periods = (last_d >> kRebaseShift) + 1;
  if (warp == 0 && Tb > 0) {
    // ---- alpha: lane r owns slot r and walks the frames upwards; the record of the next cell is fetched
    // one cell ahead, so the dependent chain of a step is shuffle -> log2-add.
    const bool act = lane < R;
    const BandStep* rp = arec + (act ? lane : 0);
    BandStep cur = *rp;
    if (!act) cur.f = kBandDone;
    rp += R;
    BandStep nxt = *rp;  // frame 1 (or the terminator written for t = T_b)
    rp += R;             // rp always points one record past nxt (reads beyond the terminator are never used)
    float* ap = als + lane;
    float last = kNegInf;
    double off = 0.0;
    for (int p = 0; p < n_periods; ++p) {
      if (lane == 0) aoff[p] = off;
#pragma unroll 4
      for (int ii = 0; ii < 16; ++ii) {
        const int d = 16 * p + ii;
        const float up_src = __shfl_up_sync(0xffffffffu, last, 1);
        const float left_src = __shfl_sync(0xffffffffu, last, (lane + cur.w) & 31);
        const BandStep cand = *rp;  // branch-free: fetched every step, consumed only when the lane advances
        const bool on = cur.f == d;
        const float val = on ? log2_add(up_src + cur.x, ((cur.w & 256) ? 0.f : left_src) + cur.y) : kNegInf;
        if (on) *ap = val;
        ap += on ? R : 0;
        rp += on ? R : 0;
        cur.x = on ? nxt.x : cur.x; cur.y = on ? nxt.y : cur.y; cur.f = on ? nxt.f : cur.f; cur.w = on ? nxt.w : cur.w;
        nxt.x = on ? cand.x : nxt.x; nxt.y = on ? cand.y : nxt.y; nxt.f = on ? cand.f : nxt.f; nxt.w = on ? cand.w : nxt.w;
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
    const bool act = lane < R;
    const BandStep* rp = brec + (Tb - 1) * R + (act ? lane : 0);
    BandStep cur = *rp;
    if (!act) cur.f = -kBandDone;
    rp -= R;
    BandStep nxt = *rp;  // frame T_b - 2, or the terminator stored as frame -1
    rp -= R;             // rp always points one record past nxt
    float* bp = bes + (Tb - 1) * R + lane;
    float last = kNegInf;
    double off = 0.0;
    for (int p = n_periods - 1; p >= 0; --p) {
      if (lane == 0) boff[p] = off;
#pragma unroll 4
      for (int ii = 15; ii >= 0; --ii) {
        const int d = 16 * p + ii;
        const float down_src = __shfl_down_sync(0xffffffffu, last, 1);
        const float right_src = __shfl_sync(0xffffffffu, last, (lane - cur.w) & 31);
        const BandStep cand = *rp;
        const bool on = cur.f == d;
        const float val = on ? log2_add(down_src + cur.x, ((cur.w & 256) ? 0.f : right_src) + cur.y) : kNegInf;
        if (on) *bp = val;
        bp -= on ? R : 0;
        rp -= on ? R : 0;
        cur.x = on ? nxt.x : cur.x; cur.y = on ? nxt.y : cur.y; cur.f = on ? nxt.f : cur.f; cur.w = on ? nxt.w : cur.w;
        nxt.x = on ? cand.x : nxt.x; nxt.y = on ? cand.y : nxt.y; nxt.f = on ? cand.f : nxt.f; nxt.w = on ? cand.w : nxt.w;
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


size_t simple_lattice_fast_workspace_bytes(int B, int S, int T) {
  const int rpl = simple_rpl(S);
  if (!rpl) return 0;
  const size_t diag_rows = (size_t)((S + T) / 32 + 1) * 32;
  const size_t diag = (size_t)B * diag_rows * 32 * rpl * sizeof(float);
  const size_t n_off = diag_rows / 16 + 2;
  return 2 * diag + (2 * (size_t)B * rpl * n_off + B) * sizeof(double) + 64;
}

static size_t simple_smem_bytes(int G, int diag_rows, int n_off) {
  const int ring = G <= 4 ? 3 : 2;
  return (size_t)2 * G * 32 * ring * 33 * sizeof(float) + (size_t)(G + 1) * (diag_rows + 2) * sizeof(float) + 8 +
         (size_t)G * n_off * sizeof(double) + (size_t)(2 * G * ring + G * (diag_rows / 32)) * 8 + 16;
}

bool simple_lattice_fast_ok(int S, int T) {
  const int g = simple_rpl(S);
  if (!g) return false;
  const int diag_rows = ((S + T) / 32 + 1) * 32;
  return simple_smem_bytes(g, diag_rows, diag_rows / 16 + 2) <= 227 * 1024;
}

int launch_simple_lattice_fast(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                               void* ws, float* logp, float* occ_px, float* occ_py, cudaStream_t stream) {
  const int rpl = simple_rpl(S);
  S2T_REQUIRE(rpl > 0, "simple_lattice_fast: S+1 = %d > 256", S + 1);
  if (B == 0) return 0;
  SimpleArgs a{};
  a.px = px; a.py = py; a.boundary = boundary;
  a.B = B; a.S = S; a.T = T;
  a.groups = rpl;
  a.rows_pad = 32 * rpl;
  a.diag_rows = ((S + T) / 32 + 1) * 32;
  const size_t diag = (size_t)B * a.diag_rows * a.rows_pad;
  a.n_off = a.diag_rows / 16 + 2;
  a.alpha_diag = (float*)ws;
  a.beta_diag = a.alpha_diag + diag;
  a.aoff = (double*)(a.beta_diag + diag + ((2 * diag) & 1));
  a.boff = a.aoff + (size_t)B * rpl * a.n_off;
  a.logp_d = a.boff + (size_t)B * rpl * a.n_off;
  a.logp = logp;
  const size_t smem = simple_smem_bytes(rpl, a.diag_rows, a.n_off);
  S2T_REQUIRE(smem <= 227 * 1024, "simple_lattice_fast: S+T = %d needs %zu B of shared memory", S + T, smem);
  {
    ProfScope prof("simple_lattice_kernel", stream);
    dim3 grid(B, occ_px ? 2 : 1);
    auto launch = [&](auto kern, int threads) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, threads, smem, stream>>>(a);
    };
    switch (rpl) {
      case 1: launch(simple_lattice_kernel<1>, 32 * (1 + kSimpleProducerWarps)); break;
      case 2: launch(simple_lattice_kernel<2>, 32 * (2 + kSimpleProducerWarps)); break;
      case 4: launch(simple_lattice_kernel<4>, 32 * (4 + kSimpleProducerWarps)); break;
      default: launch(simple_lattice_kernel<8>, 32 * (8 + kSimpleProducerWarps)); break;
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

// ---- unpruned lattice in band layout -> the systolic full-lattice kernel ---------------------------------------
// The vanilla RNN-T loss hands px / py over as (B, T, S+1) (frame-major, ranges == NULL): a full lattice.  Its
// recursion runs on simple_lattice_kernel after a tiled transpose into the k2 layout, and the occupation
// probabilities are transposed back.
namespace {

// in (B, T, R=S+1) -> px_k2 (B, S, T+1) [-inf in column T_b and beyond the symbols], py_k2 (B, S+1, T)
__global__ void __launch_bounds__(256) full_to_k2_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                         const int64_t* __restrict__ boundary, int S, int T,
                                                         float* __restrict__ px_k2, float* __restrict__ py_k2) {
  __shared__ float tx[32][33], ty[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, s0 = blockIdx.y * 32, R = S + 1;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
  const int Tb = boundary ? (int)boundary[4 * b + 3] : T;
  for (int i = ly; i < 32; i += 8) {  // rows t, columns s (s contiguous in the source)
    const int t = t0 + i, s = s0 + lx;
    const bool ok = t < T && s < R;
    tx[i][lx] = ok ? px[((int64_t)b * T + t) * R + s] : kNegInf;
    ty[i][lx] = ok ? py[((int64_t)b * T + t) * R + s] : kNegInf;
  }
  __syncthreads();
  for (int i = ly; i < 32; i += 8) {  // rows s, columns t (t contiguous in the destination)
    const int s = s0 + i, t = t0 + lx;
    if (s >= R) continue;
    if (t < T) py_k2[((int64_t)b * R + s) * T + t] = ty[lx][i];
    if (s < S && t < T) px_k2[((int64_t)b * S + s) * (T + 1) + t] = (t == Tb) ? kNegInf : tx[lx][i];
    if (s < S && t == T - 1) px_k2[((int64_t)b * S + s) * (T + 1) + T] = kNegInf;
  }
}

// occupation probabilities back: k2 layout -> (B, T, R)
__global__ void __launch_bounds__(256) k2_to_full_kernel(const float* __restrict__ ox_k2, const float* __restrict__ oy_k2,
                                                         int S, int T, float* __restrict__ occ_px,
                                                         float* __restrict__ occ_py) {
  __shared__ float tx[32][33], ty[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, s0 = blockIdx.y * 32, R = S + 1;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int i = ly; i < 32; i += 8) {  // rows s, columns t
    const int s = s0 + i, t = t0 + lx;
    tx[i][lx] = (s < S && t < T) ? ox_k2[((int64_t)b * S + s) * (T + 1) + t] : 0.f;
    ty[i][lx] = (s < R && t < T) ? oy_k2[((int64_t)b * R + s) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ly; i < 32; i += 8) {  // rows t, columns s
    const int t = t0 + i, s = s0 + lx;
    if (t >= T || s >= R) continue;
    occ_px[((int64_t)b * T + t) * R + s] = tx[lx][i];
    occ_py[((int64_t)b * T + t) * R + s] = ty[lx][i];
  }
}

size_t full_k2_floats(int B, int S, int T) {
  return 2 * ((size_t)B * S * (T + 1) + (size_t)B * (S + 1) * T);  // px, py, occ_px, occ_py in k2 layout
}

}  // namespace

bool full_lattice_fast_ok(int S, int T) { return S >= 1 && simple_lattice_fast_ok(S, T); }

size_t full_lattice_fast_workspace_bytes(int B, int S, int T) {
  if (!full_lattice_fast_ok(S, T)) return 0;
  return ((full_k2_floats(B, S, T) * sizeof(float) + 255) / 256) * 256 + simple_lattice_fast_workspace_bytes(B, S, T);
}

int launch_full_lattice_fast(const float* px, const float* py, const int64_t* boundary, int B, int S, int T, void* ws,
                             float* logp, float* occ_px, float* occ_py, cudaStream_t stream) {
  if (B == 0 || T == 0) return 0;
  float* px_k2 = (float*)ws;
  float* py_k2 = px_k2 + (size_t)B * S * (T + 1);
  float* ox_k2 = py_k2 + (size_t)B * (S + 1) * T;
  float* oy_k2 = ox_k2 + (size_t)B * S * (T + 1);
  void* ws2 = (char*)ws + ((full_k2_floats(B, S, T) * sizeof(float) + 255) / 256) * 256;
  const dim3 grid((unsigned)((T + 31) / 32), (unsigned)((S + 1 + 31) / 32), (unsigned)B);
  {
    ProfScope prof("full_lattice_transpose_kernels", stream);
    full_to_k2_kernel<<<grid, 256, 0, stream>>>(px, py, boundary, S, T, px_k2, py_k2);
  }
  if (int rc = check_launch("full_to_k2_kernel")) return rc;
  if (int rc = launch_simple_lattice_fast(px_k2, py_k2, boundary, B, S, T, ws2, logp, occ_px ? ox_k2 : nullptr,
                                          occ_px ? oy_k2 : nullptr, stream))
    return rc;
  if (occ_px) {
    ProfScope prof("full_lattice_transpose_kernels", stream);
    k2_to_full_kernel<<<grid, 256, 0, stream>>>(ox_k2, oy_k2, S, T, occ_px, occ_py);
    return check_launch("k2_to_full_kernel");
  }
  return 0;
}

static size_t band_smem_bytes(int S, int T, int R) {
  const size_t n_off = ((S + T + R) >> kRebaseShift) + 2;
  return (size_t)4 * T * R * sizeof(float) + 16 + (size_t)(2 * (T + 1) + 3) * R * 16 + (size_t)(T + 2) * sizeof(int) + 8 +
         2 * n_off * sizeof(double) + 16;
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
