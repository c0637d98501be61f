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
// The kernel is latency bound by the (S_b + T_b + 1) sequential diagonals, not
// by HBM: algorithmic traffic is 2 reads + 1 alpha write per live cell forward,
// 3 reads + 2 writes backward.
#include "lattice.cuh"

namespace s2t {
namespace {

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
    if (s > 0 && t < Tb) {
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
  int Sb, Tb;
  utt_dims(v, b, Sb, Tb);
  const float* px = v.px + (int64_t)b * v.px_bs;
  const float* py = v.py + (int64_t)b * v.py_bs;
  const int64_t* rg = v.ranges ? v.ranges + (int64_t)b * v.rg_bs : nullptr;
  float* alpha = v.alpha + (int64_t)b * v.a_bs;

  const int nd = Sb + Tb;
  float p_left = kNegInf;
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
    __syncthreads();
  }
  if (s == Sb) logp[b] = p_left;
}

struct BetaIn {
  float xv;        // px(s, t) or -inf
  float yv;        // py(s, t) or -inf
  float av;        // alpha(s, t)
  int64_t x_off;   // occ_px offset or -1
  int64_t y_off;   // occ_py offset or -1
  bool active;
};

__device__ __forceinline__ BetaIn beta_fetch(const LatticeView& v, const float* px, const float* py,
                                             const float* alpha, const int64_t* rg, int s, int d,
                                             int Sb, int Tb) {
  BetaIn in;
  in.xv = kNegInf;
  in.yv = kNegInf;
  in.av = kNegInf;
  in.x_off = -1;
  in.y_off = -1;
  int t = d - s;
  in.active = (s <= Sb) && (t >= 0) && (t <= Tb);
  if (in.active && t < Tb) {
    int r = s - sb_of(rg, v.rg_ts, t, Tb);
    if (r >= 0 && r < v.ry) {
      in.y_off = (int64_t)t * v.py_ts + (int64_t)r * v.py_rs;
      in.yv = __ldg(py + in.y_off);
      in.av = alpha[(int64_t)t * v.a_ts + (int64_t)r * v.a_rs];
      if (r < v.rx && s < Sb) {
        in.x_off = (int64_t)t * v.px_ts + (int64_t)r * v.px_rs;
        in.xv = __ldg(px + in.x_off);
      }
    }
  }
  return in;
}

__global__ void lattice_beta_kernel(LatticeView v, const float* __restrict__ logp,
                                    float* __restrict__ occ_px, float* __restrict__ occ_py) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, s = threadIdx.x, n = blockDim.x;
  float* buf[2] = {sm, sm + (n + 1)};
  int Sb, Tb;
  utt_dims(v, b, Sb, Tb);
  const float* px = v.px + (int64_t)b * v.px_bs;
  const float* py = v.py + (int64_t)b * v.py_bs;
  const float* alpha = v.alpha + (int64_t)b * v.a_bs;
  const int64_t* rg = v.ranges ? v.ranges + (int64_t)b * v.rg_bs : nullptr;
  float* ox = occ_px + (int64_t)b * v.px_bs;
  float* oy = occ_py + (int64_t)b * v.py_bs;
  const float lp = logp[b];
  const bool lp_ok = (lp - lp == 0.f);

  const int nd = Sb + Tb;
  float b_right = kNegInf;
  buf[0][s] = kNegInf;
  buf[1][s] = kNegInf;
  if (s == 0) {
    buf[0][n] = kNegInf;
    buf[1][n] = kNegInf;
  }
  __syncthreads();
  BetaIn nxt = beta_fetch(v, px, py, alpha, rg, s, nd, Sb, Tb);
  for (int d = nd; d >= 0; --d) {
    const BetaIn cur = nxt;
    if (d > 0) nxt = beta_fetch(v, px, py, alpha, rg, s, d - 1, Sb, Tb);
    float* cbuf = buf[d & 1];
    const float* nbuf = buf[(d & 1) ^ 1];
    float val = kNegInf;
    if (cur.active) {
      float bx = cur.xv + nbuf[s + 1];
      float by = cur.yv + b_right;
      val = (d == nd) ? 0.f : log_add(bx, by);  // d == nd <=> (s, t) == (S_b, T_b)
      if (lp_ok) {
        if (cur.y_off >= 0) oy[cur.y_off] = expf(cur.av + by - lp);
        if (cur.x_off >= 0) ox[cur.x_off] = expf(cur.av + bx - lp);
      }
      b_right = val;
    }
    cbuf[s] = val;
    __syncthreads();
  }
}

int threads_for(int S) { return ((S + 1 + 31) / 32) * 32; }

}  // namespace

int launch_lattice_fwd(const LatticeView& v, float* logp, cudaStream_t stream) {
  S2T_REQUIRE(v.S + 1 <= 1024, "lattice DP: S+1 = %d exceeds the 1024 symbol positions one CTA covers",
              v.S + 1);
  if (v.B == 0) return 0;
  int n = threads_for(v.S);
  lattice_alpha_kernel<<<v.B, n, 2 * n * sizeof(float), stream>>>(v, logp);
  return check_launch("lattice_alpha_kernel");
}

int launch_lattice_fwd_bwd(const LatticeView& v, float* logp, float* occ_px, float* occ_py,
                           cudaStream_t stream) {
  int rc = launch_lattice_fwd(v, logp, stream);
  if (rc) return rc;
  if (v.B == 0) return 0;
  int n = threads_for(v.S);
  lattice_beta_kernel<<<v.B, n, 2 * (n + 1) * sizeof(float), stream>>>(v, logp, occ_px, occ_py);
  return check_launch("lattice_beta_kernel");
}

}  // namespace s2t
