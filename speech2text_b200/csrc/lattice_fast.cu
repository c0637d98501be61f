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

__device__ __forceinline__ float log_add_fast(float x, float y) {
  // same cut-off as k2's LogAdd; exp / log1p through the SFU
  float mx = fmaxf(x, y), mn = fminf(x, y);
  float d = mn - mx;  // <= 0, NaN when both are -inf
  return (d >= kMinLogDiff) ? mx + log1pf(__expf(d)) : mx;
}

// ------------------------------------------------------------------------------------------------
// (1) simple lattice
// ------------------------------------------------------------------------------------------------
struct SimpleArgs {
  const float* px;  // (B, S, T+1)
  const float* py;  // (B, S+1, T)
  const int64_t* boundary;
  int B, S, T;
  int rows_pad;      // 32 * RPL
  float* alpha_diag;  // (B, S+T+1, rows_pad)
  float* beta_diag;
  double* aoff;  // (B, n_off)
  double* boff;
  int n_off;
  double* logp_d;  // (B)
  float* logp;     // (B)
};

template <int RPL>
__global__ void __launch_bounds__(160, 1) simple_lattice_kernel(SimpleArgs a) {
  constexpr int ROWS = 32 * RPL;
  constexpr int RS = ROWS + 1;  // odd stride: the producers' diagonal scatter is conflict-free
  constexpr int W = 32 * kRingBlocks;
  extern __shared__ float sm[];
  float* ringX = sm;
  float* ringY = sm + W * RS;
  uint64_t* full = reinterpret_cast<uint64_t*>(ringY + W * RS + (((W * RS * 2) & 1) ? 1 : 0));
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
      mbar_init(&full[i], 4);
      mbar_init(&empty[i], 1);
    }
    tc::fence_mbar_init();
  }
  __syncthreads();

  if (warp > 0) {
    // ---------------- producers: 4 warps, rows s = pw, pw + 4, ... ----------------
    const int pw = warp - 1;
    for (int k = 0; k < nblk; ++k) {
      const int D = is_beta ? (nblk - 1 - k) : k;  // diagonal block, processing order
      const int slot = k % kRingBlocks;
      mbar_wait(&empty[slot], ((k / kRingBlocks) & 1) ^ 1);
      const int d = 32 * D + lane;  // this lane's diagonal
      float* rx = ringX + (slot * 32 + lane) * RS;
      float* ry = ringY + (slot * 32 + lane) * RS;
#pragma unroll 4
      for (int s = pw; s <= Sb; s += 4) {
        const int t = d - s;
        float xv = kNegInf, yv = kNegInf;
        if (!is_beta) {
          // alpha: X = px(s-1, t), Y = py(s, t-1)
          if (s >= 1 && t >= 0 && t <= Tb) xv = __ldg(px + (int64_t)(s - 1) * (a.T + 1) + t);
          if (t >= 1 && t <= Tb) yv = __ldg(py + (int64_t)s * a.T + t - 1);
        } else {
          // beta: X = px(s, t), Y = py(s, t)
          if (s < Sb && t >= 0 && t <= Tb) xv = __ldg(px + (int64_t)s * (a.T + 1) + t);
          if (t >= 0 && t < Tb) yv = __ldg(py + (int64_t)s * a.T + t);
        }
        rx[s] = xv;
        ry[s] = yv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    }
    return;
  }

  // ---------------- recursion warp ----------------
  float v[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) v[j] = kNegInf;
  double off = 0.0;
  float* out = (is_beta ? a.beta_diag : a.alpha_diag) + (int64_t)b * (a.S + a.T + 1) * ROWS;
  double* offs = (is_beta ? a.boff : a.aoff) + (int64_t)b * a.n_off;

  for (int k = 0; k < nblk; ++k) {
    const int D = is_beta ? (nblk - 1 - k) : k;
    const int slot = k % kRingBlocks;
    mbar_wait(&full[slot], (k / kRingBlocks) & 1);
    for (int ii = 0; ii < 32; ++ii) {
      const int i = is_beta ? 31 - ii : ii;
      const int d = 32 * D + i;
      if (d > nd) continue;
      // offsets are piecewise constant over 16 diagonals; record the one this diagonal is stored with
      const bool period_start = is_beta ? ((d & 15) == 15 || d == nd) : ((d & 15) == 0);
      if (period_start && lane == 0) offs[d >> kRebaseShift] = off;
      const float* rx = ringX + (slot * 32 + i) * RS;
      const float* ry = ringY + (slot * 32 + i) * RS;
      float nv[RPL];
      float rot[RPL];
#pragma unroll
      for (int j = 0; j < RPL; ++j) rot[j] = __shfl_sync(0xffffffffu, v[j], is_beta ? ((lane + 1) & 31) : ((lane + 31) & 31));
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        const int s = j * 32 + lane;
        const int t = d - s;
        float nb;  // neighbour row: s-1 (alpha) / s+1 (beta), previous diagonal
        if (!is_beta) nb = (lane > 0) ? rot[j] : (j > 0 ? rot[j - 1] : kNegInf);
        else nb = (lane < 31) ? rot[j] : (j < RPL - 1 ? rot[j + 1] : kNegInf);
        const float xv = rx[s], yv = ry[s];
        float val = log_add_fast(nb + xv, v[j] + yv);
        const bool start = is_beta ? (s == Sb && t == Tb) : (d == 0 && s == 0);
        if (start) val = 0.f;
        if (s > Sb || t < 0 || t > Tb) val = kNegInf;
        nv[j] = val;
        out[(int64_t)d * ROWS + s] = val;
      }
#pragma unroll
      for (int j = 0; j < RPL; ++j) v[j] = nv[j];
      const bool period_end = is_beta ? ((d & 15) == 0) : ((d & 15) == 15);
      if (period_end) {
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
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
  if (!is_beta) {
    // log P(y|x) = alpha(S_b, T_b), held by lane S_b % 32, register S_b / 32 (diagonal nd was the last)
    float mine = kNegInf;
#pragma unroll
    for (int j = 0; j < RPL; ++j)
      if (j == Sb / 32) mine = v[j];
    const float fin = __shfl_sync(0xffffffffu, mine, Sb & 31);
    if (lane == 0) {
      const double lp = (double)fin + off;
      a.logp_d[b] = lp;
      a.logp[b] = (float)lp;
    }
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
  const float* ad = a.alpha_diag + (int64_t)b * (a.S + a.T + 1) * ROWS;
  const float* bd = a.beta_diag + (int64_t)b * (a.S + a.T + 1) * ROWS;
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
      const float cst = (float)(aoff[d >> kRebaseShift] + boff[(d + 1 <= nd ? d + 1 : nd) >> kRebaseShift] - lp);
      const float av = sA[rs + tx][rs];
      if (t < Tb) {
        const float yv = __ldg(a.py + ((int64_t)b * (a.S + 1) + s) * a.T + t);
        oy = __expf(av + yv + sB[rs + tx][rs] + cst);
      }
      if (s < Sb) {
        const float xv = __ldg(a.px + ((int64_t)b * a.S + s) * (a.T + 1) + t);
        ox = __expf(av + xv + sB[rs + tx][rs + 1] + cst);
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
  float* pxs = sm;
  float* pys = pxs + T * R;
  float* als = pys + T * R;
  float* bes = als + T * R;
  int* sbs = reinterpret_cast<int*>(bes + T * R);  // sb[t], T + 1 entries (sb[Tb] = sb[Tb-1])
  double* offs = reinterpret_cast<double*>(sbs + ((T + 2) & ~1));  // aoff | boff, n_off each
  const int n_off = ((a.S + T + R) >> kRebaseShift) + 2;
  double* aoff = offs;
  double* boff = offs + n_off;
  __shared__ double s_logp;

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
    pxs[i] = __ldg(gpx + i);
    pys[i] = __ldg(gpy + i);
  }
  for (int t = threadIdx.x; t <= Tb; t += blockDim.x) {
    const int tt = min(t, max(Tb - 1, 0));
    sbs[t] = grg ? (int)grg[(int64_t)tt * R] : 0;
  }
  if (threadIdx.x == 0) s_logp = -INFINITY;
  __syncthreads();

  const int last_d = (Tb > 0) ? (Tb - 1 + sbs[Tb - 1] + R - 1) : -1;  // last diagonal that holds a band cell
  if (warp == 0 && Tb > 0) {
    // ---- alpha: lane r owns slot r, frames ascending ----
    int t = 0;
    float last = kNegInf;
    double off = 0.0;
    for (int d = 0; d <= last_d; ++d) {
      if ((d & 15) == 0 && lane == 0) aoff[d >> kRebaseShift] = off;
      const bool act = (lane < R) && (t < Tb) && (t + sbs[t] + lane == d);
      const int delta = (act && t > 0) ? (sbs[t] - sbs[t - 1]) : 0;
      const float up_src = __shfl_up_sync(0xffffffffu, last, 1);
      const float left_src = __shfl_sync(0xffffffffu, last, (lane + delta) & 31);
      float val = kNegInf;
      if (act) {
        const int s = sbs[t] + lane;
        if (s <= Sb) {
          const float up = (lane > 0) ? up_src + pxs[t * R + lane - 1] : kNegInf;
          const float left = (t > 0 && lane + delta < R) ? left_src + pys[(t - 1) * R + lane + delta] : kNegInf;
          val = (t == 0 && s == 0) ? 0.f : log_add_fast(up, left);
        }
        als[t * R + lane] = val;
        ++t;
      }
      last = val;
      if ((d & 15) == 15) {
        const float m = warp_max(last);
        // every lane's `last` is from this diagonal or -inf; older cells are already in als[]
        if (m - m == 0.f) {
          last -= m;
          off += (double)m;
        }
      }
    }
    // log P = alpha(T_b-1, r*) + py(T_b-1, r*), r* = S_b - sb[T_b-1]
    if (lane == 0) {
      const int rs = Sb - sbs[Tb - 1];
      double lp = -INFINITY;
      if (rs >= 0 && rs < R) {
        const int dd = Tb - 1 + sbs[Tb - 1] + rs;
        // als[] values are relative to the offset in force on their own diagonal
        lp = (double)als[(Tb - 1) * R + rs] + (double)pys[(Tb - 1) * R + rs];
        (void)dd;
        s_logp = lp;  // completed after the barrier with aoff[dd >> 4], the offset of that diagonal
      }
    }
  } else if (warp == 1 && Tb > 0) {
    // ---- beta: lane r owns slot r, frames descending ----
    int t = Tb - 1;
    float last = kNegInf;
    double off = 0.0;
    for (int d = last_d; d >= 0; --d) {
      if (((d & 15) == 15 || d == last_d) && lane == 0) boff[d >> kRebaseShift] = off;
      const bool act = (lane < R) && (t >= 0) && (t + sbs[t] + lane == d);
      const int delta = (act && t + 1 < Tb) ? (sbs[t + 1] - sbs[t]) : 0;
      const float down_src = __shfl_down_sync(0xffffffffu, last, 1);
      const float right_src = __shfl_sync(0xffffffffu, last, (lane - delta) & 31);
      float val = kNegInf;
      if (act) {
        const int s = sbs[t] + lane;
        if (s <= Sb) {
          // symbol: to (s+1, t) = slot lane+1 of the same frame
          const float bx = (lane + 1 < R && s < Sb) ? pxs[t * R + lane] + down_src : kNegInf;
          float by;
          if (t == Tb - 1) by = (s == Sb) ? pys[t * R + lane] : kNegInf;  // beta(s, T_b) = [s == S_b]
          else by = (lane - delta >= 0) ? pys[t * R + lane] + right_src : kNegInf;
          val = log_add_fast(bx, by);
        }
        bes[t * R + lane] = val;
        --t;
      }
      last = val;
      if ((d & 15) == 0) {
        const float m = warp_max(last);
        if (m - m == 0.f) {
          last -= m;
          off += (double)m;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && Tb > 0) {
    const int rs = Sb - sbs[Tb - 1];
    if (rs >= 0 && rs < R) {
      const int dd = Tb - 1 + sbs[Tb - 1] + rs;
      s_logp = s_logp + aoff[dd >> kRebaseShift];
    }
    a.logp[b] = (float)s_logp;
  } else if (threadIdx.x == 0) {
    a.logp[b] = (Sb == 0) ? 0.f : -INFINITY;  // no frames: only the empty transcript is possible
  }
  __syncthreads();
  const double lp = s_logp;
  const bool lp_ok = (lp - lp == 0.0) && Tb > 0;
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
        // blank: to (s, t+1)
        if (t == Tb - 1) {
          if (s == Sb) vy = __expf((float)((double)av + (double)pys[i] + ao - lp));
        } else {
          const int r2 = s - sbs[t + 1];
          if (r2 >= 0 && r2 < R) {
            const float cst = (float)(ao + boff[(d + 1) >> kRebaseShift] - lp);
            vy = __expf(av + pys[i] + bes[(t + 1) * R + r2] + cst);
          }
        }
        // symbol: to (s+1, t)
        if (r + 1 < R && s < Sb) {
          const float cst = (float)(ao + boff[(d + 1) >> kRebaseShift] - lp);
          vx = __expf(av + pxs[i] + bes[t * R + r + 1] + cst);
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
  const size_t diag = (size_t)B * (S + T + 1) * 32 * rpl * sizeof(float);
  const size_t n_off = ((S + T) >> kRebaseShift) + 2;
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
  const size_t diag = (size_t)B * (S + T + 1) * a.rows_pad;
  a.n_off = ((S + T) >> kRebaseShift) + 2;
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
      kern<<<grid, 160, smem, stream>>>(a);
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
  return (size_t)4 * T * R * sizeof(float) + (size_t)((T + 2) & ~1) * sizeof(int) + 2 * n_off * sizeof(double) + 16;
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
