// The "simple" (trivial-joiner) lattice loss of the pruned RNN-T recipe:
// log-probs px/py of joiner(am, lm) = am + lm, normalised over the vocabulary.
//
// Replaces k2.rnnt_loss_smoothed(..., return_grad=True) as called from
// /root/reference/model/joiner/joiner.py:100-110 (k2 semantics: SURVEY.md A.1-A.3,
// gradients A.7).  Pipeline (all fp32):
//   1. row_max_kernel       am_max (B,T), lm_max (B,S+1)
//   2. normaliser contraction  acc[b,s,t] = sum_c exp(lm-lm_max) * exp(am-am_max)   (sgemm_kernel,
//      operands exponentiated on the fly) with the px/py emit fused in the epilogue:
//        nrm = log(acc + tiny) + lm_max + am_max
//        py[b,s,t] = am[b,t,0] + lm[b,s,0] - nrm
//        px[b,s,t] = am[b,t,sym] + lm[b,s,sym] - nrm ; px[b,s,T_b] = px[b,s,T] = -inf
//   3. lattice DP (lattice.cu) -> logp, occupation probabilities px_grad / py_grad
// Backward (given d loss / d logp[b]):
//   W[b,s,t] = coef_b * (px_grad + py_grad)[s,t] * exp(am_max + lm_max - nrm)
//   d_am = -exp(am-am_max) * (W^T . exp(lm-lm_max))  + one-hot scatter of the occupations
//   d_lm = -exp(lm-lm_max) * (W   . exp(am-am_max))  + one-hot scatter
#include "lattice.cuh"
#include "sgemm.cuh"

namespace s2t {
namespace {

// one warp per row: max over V
__global__ void row_max_kernel(const float* __restrict__ x, int64_t rows, int V,
                               float* __restrict__ out) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x % 32;
  const float* p = x + row * V;
  float m = kNegInf;
  for (int c = lane; c < V; c += 32) m = fmaxf(m, __ldg(p + c));
  m = warp_max(m);
  if (lane == 0) out[row] = m;
}

// A operand: exp(x[b, idx, k] - max[b, idx]), k contiguous
struct ExpRowOperand {
  const float* x;
  const float* mx;
  int rows_per_batch, V;
  struct Row { const float* p; float m; };
  __device__ Row row(int batch, int idx) const {
    int64_t r = (int64_t)batch * rows_per_batch + idx;
    return Row{x + r * V, mx[r]};
  }
  __device__ float at(const Row& r, int k) const { return expf(__ldg(r.p + k) - r.m); }
};

// B operand with the *class* index as the row: element(n = c, k = row) = exp(x[b,k,c]-max[b,k])
struct ExpColOperand {
  const float* x;
  const float* mx;
  int rows_per_batch, V;
  struct Row { const float* p; const float* m; };
  __device__ Row row(int batch, int c) const {
    int64_t r0 = (int64_t)batch * rows_per_batch;
    return Row{x + r0 * V + c, mx + r0};
  }
  __device__ float at(const Row& r, int k) const { return expf(__ldg(r.p + (int64_t)k * V) - __ldg(r.m + k)); }
};

struct SimpleEmitEpilogue {
  const float* am;
  const float* lm;
  const float* am_max;
  const float* lm_max;
  const int64_t* sym;       // (B, S)
  const int64_t* boundary;  // (B, 4)
  int T, S, V, blank;
  float* px;   // (B, S, T+1)
  float* py;   // (B, S+1, T)
  float* nrm;  // (B, S+1, T)
  __device__ void operator()(int b, int s, int t, float acc) const {
    const float* am_row = am + ((int64_t)b * T + t) * V;
    const float* lm_row = lm + ((int64_t)b * (S + 1) + s) * V;
    float n = logf(acc + FLT_MIN) + lm_max[(int64_t)b * (S + 1) + s] + am_max[(int64_t)b * T + t];
    int64_t o = ((int64_t)b * (S + 1) + s) * T + t;
    nrm[o] = n;
    py[o] = __ldg(am_row + blank) + __ldg(lm_row + blank) - n;
    if (s < S) {
      int c = (int)sym[(int64_t)b * S + s];
      const int Tb = boundary ? (int)boundary[4 * b + 3] : T;
      float* px_row = px + ((int64_t)b * S + s) * (T + 1);
      float v = __ldg(am_row + c) + __ldg(lm_row + c) - n;
      px_row[t] = (t == Tb) ? kNegInf : v;
      if (t == T - 1) px_row[T] = kNegInf;
    }
  }
};

// W[b,s,t] = coef_b * (occ_px + occ_py) * exp(am_max + lm_max - nrm)
__global__ void simple_w_kernel(const float* __restrict__ occ_px, const float* __restrict__ occ_py,
                                const float* __restrict__ nrm, const float* __restrict__ am_max,
                                const float* __restrict__ lm_max, const float* __restrict__ coef,
                                int B, int S, int T, float* __restrict__ w) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * (S + 1) * T;
  if (i >= total) return;
  int t = (int)(i % T);
  int64_t bs = i / T;
  int s = (int)(bs % (S + 1));
  int b = (int)(bs / (S + 1));
  float g = occ_py[i];
  if (s < S) g += occ_px[((int64_t)b * S + s) * (T + 1) + t];
  float e = expf(am_max[(int64_t)b * T + t] + lm_max[bs] - nrm[i]);
  w[i] = (g == 0.f) ? 0.f : coef[b] * g * e;
}

// A operand for d_am: element(m = t, k = s) = W[b,s,t]   (m contiguous)
// A operand for d_lm: element(m = s, k = t) = W[b,s,t]   (k contiguous)
struct GradAmEpilogue {  // d_am[b,t,c] = -exp(am - am_max) * acc
  const float* am;
  const float* am_max;
  int T, V;
  float* d_am;
  __device__ void operator()(int b, int t, int c, float acc) const {
    int64_t r = (int64_t)b * T + t;
    d_am[r * V + c] = -expf(__ldg(am + r * V + c) - am_max[r]) * acc;
  }
};

// one-hot terms of d_am: thread per (b, s, t), t fastest (coalesced occupation reads), scattered atomics
__global__ void simple_scatter_am_kernel(const float* __restrict__ occ_px, const float* __restrict__ occ_py,
                                         const int64_t* __restrict__ sym, const float* __restrict__ coef,
                                         int B, int S, int T, int V, int blank, float* __restrict__ d_am) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * (S + 1) * T;
  if (i >= total) return;
  const int t = (int)(i % T);
  const int64_t bs = i / T;
  const int s = (int)(bs % (S + 1));
  const int b = (int)(bs / (S + 1));
  const float cf = coef[b];
  float* row = d_am + ((int64_t)b * T + t) * V;
  const float oy = occ_py[i];
  if (oy != 0.f) atomicAdd(row + blank, cf * oy);
  if (s < S) {
    const float ox = occ_px[((int64_t)b * S + s) * (T + 1) + t];
    if (ox != 0.f) atomicAdd(row + sym[(int64_t)b * S + s], cf * ox);
  }
}

// one warp per (b, s): add the one-hot terms to d_lm
__global__ void simple_scatter_lm_kernel(const float* __restrict__ occ_px, const float* __restrict__ occ_py,
                                         const int64_t* __restrict__ sym, const float* __restrict__ coef,
                                         int B, int S, int T, int V, int blank, float* __restrict__ d_lm) {
  int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (w >= (int64_t)B * (S + 1)) return;
  const int lane = threadIdx.x % 32;
  int s = (int)(w % (S + 1)), b = (int)(w / (S + 1));
  float xs = 0.f, ys = 0.f;
  const float* py_row = occ_py + w * T;
  for (int t = lane; t < T; t += 32) ys += py_row[t];
  if (s < S) {
    const float* px_row = occ_px + ((int64_t)b * S + s) * (T + 1);
    for (int t = lane; t < T; t += 32) xs += px_row[t];
  }
  xs = warp_sum(xs);
  ys = warp_sum(ys);
  if (lane == 0) {
    float* row = d_lm + w * V;
    const float cf = coef[b];
    if (s < S) row[sym[(int64_t)b * S + s]] += cf * xs;
    row[blank] += cf * ys;
  }
}

}  // namespace

int simple_scatter_onehot(const float* occ_px, const float* occ_py, const int64_t* sym, const float* coef, int B,
                          int S, int T, int V, int blank, float* d_am, float* d_lm, cudaStream_t stream);

int simple_logprobs(const float* am, const float* lm, const int64_t* sym, const int64_t* boundary,
                    int B, int T, int S, int V, int blank, float* am_max, float* lm_max, float* px,
                    float* py, float* nrm, cudaStream_t stream) {
  const int wpb = 8;
  int64_t rows_am = (int64_t)B * T, rows_lm = (int64_t)B * (S + 1);
  {
    ProfScope prof("row_max_kernel", stream, 2);
    row_max_kernel<<<(unsigned)((rows_am + wpb - 1) / wpb), wpb * 32, 0, stream>>>(am, rows_am, V, am_max);
    row_max_kernel<<<(unsigned)((rows_lm + wpb - 1) / wpb), wpb * 32, 0, stream>>>(lm, rows_lm, V, lm_max);
  }
  if (int rc = check_launch("row_max_kernel")) return rc;
  ExpRowOperand a{lm, lm_max, S + 1, V};
  ExpRowOperand bop{am, am_max, T, V};
  SimpleEmitEpilogue ep{am, lm, am_max, lm_max, sym, boundary, T, S, V, blank, px, py, nrm};
  return launch_sgemm<true, true>(B, S + 1, T, V, 1, a, bop, ep, stream, "simple_normaliser_gemm");
}

int simple_backward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                    const float* lm_max, const float* nrm, const float* occ_px, const float* occ_py,
                    const float* coef, int B, int T, int S, int V, int blank, float* wbuf, float* d_am,
                    float* d_lm, cudaStream_t stream) {
  int64_t total = (int64_t)B * (S + 1) * T;
  if (total == 0) return 0;
  {
    ProfScope prof("simple_w_kernel", stream);
    simple_w_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(occ_px, occ_py, nrm, am_max, lm_max,
                                                                          coef, B, S, T, wbuf);
  }
  if (int rc = check_launch("simple_w_kernel")) return rc;
  // d_am: M = T (m = t), N = V, K = S+1
  {
    StridedOperand a{wbuf, (int64_t)(S + 1) * T, 1, T};  // (t, s) -> W[b, s, t]
    ExpColOperand bop{lm, lm_max, S + 1, V};
    GradAmEpilogue ep{am, am_max, T, V, d_am};
    if (int rc = launch_sgemm<false, false>(B, T, V, S + 1, 1, a, bop, ep, stream, "simple_d_am_gemm")) return rc;
  }
  // d_lm: M = S+1 (m = s), N = V, K = T
  {
    StridedOperand a{wbuf, (int64_t)(S + 1) * T, T, 1};  // (s, t) -> W[b, s, t]
    ExpColOperand bop{am, am_max, T, V};
    GradAmEpilogue ep{lm, lm_max, S + 1, V, d_lm};
    if (int rc = launch_sgemm<true, false>(B, S + 1, V, T, 1, a, bop, ep, stream, "simple_d_lm_gemm")) return rc;
  }
  return simple_scatter_onehot(occ_px, occ_py, sym, coef, B, S, T, V, blank, d_am, d_lm, stream);
}

// adds the one-hot terms of the occupation probabilities to d_am and d_lm
int simple_scatter_onehot(const float* occ_px, const float* occ_py, const int64_t* sym, const float* coef, int B,
                          int S, int T, int V, int blank, float* d_am, float* d_lm, cudaStream_t stream) {
  const int64_t total = (int64_t)B * (S + 1) * T;
  if (total == 0) return 0;
  ProfScope prof2("simple_scatter_kernels", stream, 2);
  simple_scatter_am_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(occ_px, occ_py, sym, coef, B, S,
                                                                                 T, V, blank, d_am);
  int64_t nbs = (int64_t)B * (S + 1);
  simple_scatter_lm_kernel<<<(unsigned)((nbs + 7) / 8), 256, 0, stream>>>(occ_px, occ_py, sym, coef, B, S, T,
                                                                           V, blank, d_lm);
  return check_launch("simple_scatter kernels");
}

// the same scatter with separate per-utterance factors for the am and the lm side (either may be null: skipped)
int simple_scatter_onehot_split(const float* occ_px, const float* occ_py, const int64_t* sym, const float* coef_am,
                                const float* coef_lm, int B, int S, int T, int V, int blank, float* d_am, float* d_lm,
                                cudaStream_t stream) {
  const int64_t total = (int64_t)B * (S + 1) * T;
  if (total == 0) return 0;
  ProfScope prof2("simple_scatter_kernels", stream, 2);
  if (coef_am)
    simple_scatter_am_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(occ_px, occ_py, sym, coef_am, B, S, T, V,
                                                                                   blank, d_am);
  const int64_t nbs = (int64_t)B * (S + 1);
  if (coef_lm)
    simple_scatter_lm_kernel<<<(unsigned)((nbs + 7) / 8), 256, 0, stream>>>(occ_px, occ_py, sym, coef_lm, B, S, T, V, blank,
                                                                             d_lm);
  return check_launch("simple_scatter kernels");
}

}  // namespace s2t
