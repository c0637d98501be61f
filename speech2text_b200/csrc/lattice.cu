// Anti-diagonal wavefront DP over the RNN-T lattice: alpha forward, then beta
// backward fused with the occupation probabilities (k2's px_grad / py_grad).
//
// Replaces, for the reference's hot path, k2's mutual_information_forward /
// mutual_information_backward (reached from /root/reference/model/joiner/joiner.py:100-110
// and /root/reference/model/loss/pruned_rnnt_loss.py:39-48; semantics in SURVEY.md A.2)
// and the alpha/beta part of torchaudio's rnnt_loss
// (/root/reference/model/loss/rnnt_loss.py:42-44).
//
// One CTA per utterance, one thread per symbol position s.  On diagonal d the
// thread owns cell (s, t = d - s); the neighbour value p(s-1, t) comes from a
// double-buffered shared-memory line written on diagonal d-1, p(s, t-1) stays in
// a register.  Log-probs for diagonal d+1 are fetched while diagonal d is being
// combined (their addresses do not depend on the recursion), so the dependent
// chain per step is: LDS -> FADD -> log-add-exp -> STS -> barrier.
//
// Precision: log P(y|x) reaches -3000 at BASELINE config 3, where one fp32 ulp is
// 2.4e-4 -- too coarse for occupation probabilities exp(alpha + p + beta - logP)
// that must be good to 1e-4.  Both passes therefore keep their running values
// relative to a per-CTA offset that is re-based to the diagonal maximum every
// kRebase diagonals (block max + two extra barriers, amortised); the offsets are
// accumulated in fp64 and recorded per diagonal, so the stored alpha/beta stay
// O(10) and the only fp64 arithmetic is one add per diagonal.
//
// The kernel is latency bound by the (S_b + T_b + 1) sequential diagonals, not
// by HBM: algorithmic traffic is 2 reads + 1 alpha write per live cell forward,
// 3 reads + 2 writes backward.
#include "lattice.cuh"

namespace s2t {
namespace {

constexpr int kRebase = 8;  // power of two

__device__ __forceinline__ void utt_dims(const LatticeView& v, int b, int& Sb, int& Tb) {
  if (v.boundary) {
    Sb = (int)v.boundary[4 * b + 2];
    Tb = (int)v.boundary[4 * b + 3];
  } else {
    Sb = v.S;
    Tb = v.T;
  }
  Sb = min(max(Sb, 0), v.S);
  Tb = min(max(Tb, 0), v.T);
}

__device__ __forceinline__ int sb_of(const int64_t* rg, int64_t rg_ts, int t, int Tb) {
  if (rg == nullptr) return 0;
  int tt = max(min(t, Tb - 1), 0);
  return (int)rg[(int64_t)tt * rg_ts];
}

// is px(., t) defined for this frame?  (k2 layout: also the column t == T_b, where k2 put -inf)
__device__ __forceinline__ bool px_frame_ok(const LatticeView& v, int t, int Tb) {
  return t < Tb || (v.px_at_tb && t == Tb);
}

// max over the CTA of v (all threads get the result); `red` holds >= 32 floats
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float m = red[0];
  for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
  return m;
}

struct AlphaIn {
  float xv;       // px(s-1, t) or -inf
  float yv;       // py(s, t-1) or -inf
  int64_t a_off;  // alpha offset of (s, t) or -1 when the cell is not stored
  bool active;
};

__device__ __forceinline__ AlphaIn alpha_fetch(const LatticeView& v, const float* px, const float* py,
                                               const int64_t* rg, int s, int d, int Sb, int Tb) {
  AlphaIn in;
  in.xv = kNegInf;
  in.yv = kNegInf;
  in.a_off = -1;
  int t = d - s;
  in.active = (s <= Sb) && (t >= 0) && (t <= Tb);
  if (in.active) {
    int r = s - sb_of(rg, v.rg_ts, t, Tb);
    if (s > 0 && px_frame_ok(v, t, Tb)) {
      int rr = r - 1;
      if (rr >= 0 && rr < v.rx) in.xv = __ldg(px + (int64_t)t * v.px_ts + (int64_t)rr * v.px_rs);
    }
    if (t > 0) {
      int r2 = s - sb_of(rg, v.rg_ts, t - 1, Tb);
      if (r2 >= 0 && r2 < v.ry) in.yv = __ldg(py + (int64_t)(t - 1) * v.py_ts + (int64_t)r2 * v.py_rs);
    }
    if (r >= 0 && r < v.ry) in.a_off = (int64_t)t * v.a_ts + (int64_t)r * v.a_rs;
  }
  return in;
}

__global__ void lattice_alpha_kernel(LatticeView v, float* __restrict__ logp) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, s = threadIdx.x, n = blockDim.x;
  float* buf[2] = {sm, sm + n};
  float* red = sm + 2 * n;
  int Sb, Tb;
  utt_dims(v, b, Sb, Tb);
  const float* px = v.px + (int64_t)b * v.px_bs;
  const float* py = v.py + (int64_t)b * v.py_bs;
  const int64_t* rg = v.ranges ? v.ranges + (int64_t)b * v.rg_bs : nullptr;
  float* alpha = v.alpha + (int64_t)b * v.a_bs;
  double* aoff = v.aoff + (int64_t)b * (v.S + v.T + 2);

  const int nd = Sb + Tb;
  float p_left = kNegInf;
  double off = 0.0;  // true alpha = stored alpha + off (identical in every thread)
  buf[1][s] = kNegInf;  // "previous diagonal" of d = 0
  __syncthreads();
  AlphaIn nxt = alpha_fetch(v, px, py, rg, s, 0, Sb, Tb);
  for (int d = 0; d <= nd; ++d) {
    const AlphaIn cur = nxt;
    if (d < nd) nxt = alpha_fetch(v, px, py, rg, s, d + 1, Sb, Tb);
    float* cbuf = buf[d & 1];
    const float* pbuf = buf[(d & 1) ^ 1];
    float val = kNegInf;
    if (cur.active) {
      float up = pbuf[max(s - 1, 0)] + cur.xv;
      float left = p_left + cur.yv;
      val = (d == 0) ? 0.f : log_add(up, left);
      if (cur.a_off >= 0) {
        alpha[cur.a_off] = val;
      } else {
        val = kNegInf;  // outside the band: no outgoing transition exists
      }
      p_left = val;
    }
    cbuf[s] = val;
    if (s == 0) aoff[d] = off;
    __syncthreads();
    if ((d & (kRebase - 1)) == kRebase - 1 && d < nd) {
      const float m = block_max(val, red);
      if (m - m == 0.f) {  // finite (uniform across the CTA)
        p_left -= m;
        cbuf[s] = val - m;
        off += (double)m;
        __syncthreads();
      }
    }
  }
  if (s == Sb) {
    const double lp = (double)p_left + off;
    v.logp_d[b] = lp;
    logp[b] = (float)lp;
  }
}

struct BetaIn {
  float xv;        // px(s, t) or -inf
  float yv;        // py(s, t) or -inf
  float av;        // alpha(s, t), relative to aoff[d]
  int64_t x_off;   // occ_px offset or -1
  int64_t y_off;   // occ_py offset or -1
  double aoff;     // offset of the alpha diagonal d
  bool active;
};

__device__ __forceinline__ BetaIn beta_fetch(const LatticeView& v, const float* px, const float* py,
                                             const float* alpha, const double* aoff, const int64_t* rg,
                                             int s, int d, int Sb, int Tb) {
  BetaIn in;
  in.xv = kNegInf;
  in.yv = kNegInf;
  in.av = kNegInf;
  in.x_off = -1;
  in.y_off = -1;
  in.aoff = aoff[d];
  int t = d - s;
  in.active = (s <= Sb) && (t >= 0) && (t <= Tb);
  if (in.active) {
    int r = s - sb_of(rg, v.rg_ts, t, Tb);
    if (r >= 0 && r < v.ry) {
      if (t < Tb) {
        in.y_off = (int64_t)t * v.py_ts + (int64_t)r * v.py_rs;
        in.yv = __ldg(py + in.y_off);
      }
      if (r < v.rx && s < Sb && px_frame_ok(v, t, Tb)) {
        in.x_off = (int64_t)t * v.px_ts + (int64_t)r * v.px_rs;
        in.xv = __ldg(px + in.x_off);
      }
      if (in.y_off >= 0 || in.x_off >= 0) in.av = alpha[(int64_t)t * v.a_ts + (int64_t)r * v.a_rs];
    }
  }
  return in;
}

__global__ void lattice_beta_kernel(LatticeView v, float* __restrict__ occ_px, float* __restrict__ occ_py) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, s = threadIdx.x, n = blockDim.x;
  float* buf[2] = {sm, sm + (n + 1)};
  float* red = sm + 2 * (n + 1);
  int Sb, Tb;
  utt_dims(v, b, Sb, Tb);
  const float* px = v.px + (int64_t)b * v.px_bs;
  const float* py = v.py + (int64_t)b * v.py_bs;
  const float* alpha = v.alpha + (int64_t)b * v.a_bs;
  const double* aoff = v.aoff + (int64_t)b * (v.S + v.T + 2);
  const int64_t* rg = v.ranges ? v.ranges + (int64_t)b * v.rg_bs : nullptr;
  float* ox = occ_px + (int64_t)b * v.px_bs;
  float* oy = occ_py + (int64_t)b * v.py_bs;
  const double lp = v.logp_d[b];
  const bool lp_ok = (lp - lp == 0.0);

  const int nd = Sb + Tb;
  float b_right = kNegInf;
  double off = 0.0;  // true beta = stored beta + off
  buf[0][s] = kNegInf;
  buf[1][s] = kNegInf;
  if (s == 0) {
    buf[0][n] = kNegInf;
    buf[1][n] = kNegInf;
  }
  __syncthreads();
  BetaIn nxt = beta_fetch(v, px, py, alpha, aoff, rg, s, nd, Sb, Tb);
  for (int d = nd; d >= 0; --d) {
    const BetaIn cur = nxt;
    if (d > 0) nxt = beta_fetch(v, px, py, alpha, aoff, rg, s, d - 1, Sb, Tb);
    float* cbuf = buf[d & 1];
    const float* nbuf = buf[(d & 1) ^ 1];
    float val = kNegInf;
    if (cur.active) {
      // alpha(s,t) + p + beta(next) - logP, with the two running offsets folded in fp64
      const float cst = (float)(cur.aoff + off - lp);
      float bx = cur.xv + nbuf[s + 1];
      float by = cur.yv + b_right;
      val = (d == nd) ? 0.f : log_add(bx, by);  // d == nd <=> (s, t) == (S_b, T_b)
      if (lp_ok) {
        if (cur.y_off >= 0) oy[cur.y_off] = expf(cur.av + by + cst);
        if (cur.x_off >= 0) ox[cur.x_off] = expf(cur.av + bx + cst);
      }
      b_right = val;
    }
    cbuf[s] = val;
    __syncthreads();
    if ((d & (kRebase - 1)) == 0 && d > 0) {
      const float m = block_max(val, red);
      if (m - m == 0.f) {
        b_right -= m;
        cbuf[s] = val - m;
        off += (double)m;
        __syncthreads();
      }
    }
  }
}

int threads_for(int S) { return ((S + 1 + 31) / 32) * 32; }

}  // namespace

size_t lattice_workspace_bytes(int B, int S, int T, int slots) {
  size_t alpha = (size_t)B * (T + 1) * slots * sizeof(float);
  alpha = (alpha + 15) / 16 * 16;
  return alpha + ((size_t)B * (S + T + 2) + B) * sizeof(double) + 16;
}

void lattice_carve_workspace(LatticeView& v, void* ws, int slots) {
  size_t alpha = (size_t)v.B * (v.T + 1) * slots * sizeof(float);
  alpha = (alpha + 15) / 16 * 16;
  v.alpha = (float*)ws;
  v.aoff = (double*)((char*)ws + alpha);
  v.logp_d = v.aoff + (size_t)v.B * (v.S + v.T + 2);
}

int launch_lattice_fwd(const LatticeView& v, float* logp, cudaStream_t stream) {
  S2T_REQUIRE(v.S + 1 <= 1024, "lattice DP: S+1 = %d exceeds the 1024 symbol positions one CTA covers",
              v.S + 1);
  if (v.B == 0) return 0;
  int n = threads_for(v.S);
  ProfScope prof("lattice_alpha_kernel", stream);
  lattice_alpha_kernel<<<v.B, n, (2 * n + 32) * sizeof(float), stream>>>(v, logp);
  return check_launch("lattice_alpha_kernel");
}

int launch_lattice_fwd_bwd(const LatticeView& v, float* logp, float* occ_px, float* occ_py,
                           cudaStream_t stream) {
  int rc = launch_lattice_fwd(v, logp, stream);
  if (rc) return rc;
  if (v.B == 0) return 0;
  int n = threads_for(v.S);
  ProfScope prof("lattice_beta_kernel", stream);
  lattice_beta_kernel<<<v.B, n, (2 * (n + 1) + 32) * sizeof(float), stream>>>(v, occ_px, occ_py);
  return check_launch("lattice_beta_kernel");
}

}  // namespace s2t
