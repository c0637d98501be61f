// lm-only / am-only interpolation of the simple loss (k2.get_rnnt_logprobs_smoothed with non-zero
// lm_only_scale / am_only_scale, reached from /root/reference/model/joiner/joiner.py:100-110 through
// JoinerConfig.lm_scale / am_scale; semantics: SURVEY.md A.1).
//
//   px_i = c0 px + l_s (lm[sym] - lmN) + a_s (am[sym] + log u[sym] - amN)        c0 = 1 - l_s - a_s
//   py_i = c0 py + l_s (lm[0]   - lmN) + a_s (am[0]   + log u[0]   - amN)
//   lmN[b,s] = logsumexp_c lm[b,s,c]      u[c] = mean_{b,s} softmax(lm[b,s])[c] + tiny   (ALL rows, padding included)
//   amN[b,t] = log sum_c exp(am[b,t,c]) u[c]
//
// The plain px / py / nrm come from the normaliser kernels; this file mixes them in place (forward) and adds the
// gradient terms that flow through lmN, amN and u (backward).  Everything here is row-wise work over V: one warp
// per row, block-level partial sums for the two vocabulary-sized accumulators.  A scale that is exactly 0 drops
// its term (k2 substitutes 1e-20, which vanishes in fp32).
#include "common.cuh"

namespace s2t {
namespace {

constexpr int kWarpsPerBlock = 8;

struct SmoothWs {
  float *lmN, *amN, *usum, *u, *logu, *dLdu, *H, *A, *Bt;
};

SmoothWs carve(void* ws, int B, int T, int S, int V) {
  SmoothWs w;
  float* p = (float*)ws;
  w.lmN = p; p += (size_t)B * (S + 1);
  w.amN = p; p += (size_t)B * T;
  w.usum = p; p += V;
  w.u = p; p += V;
  w.logu = p; p += V;
  w.dLdu = p; p += V;
  w.H = p; p += V;
  w.A = p; p += (size_t)B * (S + 1);
  w.Bt = p; p += (size_t)B * T;
  return w;
}

// lmN per row and the sum over rows of softmax(lm row) (block partial in shared memory, then atomics)
__global__ void __launch_bounds__(32 * kWarpsPerBlock) lm_stats_kernel(const float* __restrict__ lm,
                                                                       const float* __restrict__ lm_max, int64_t rows,
                                                                       int V, float* __restrict__ lmN,
                                                                       float* __restrict__ usum) {
  extern __shared__ float part[];  // V
  for (int c = threadIdx.x; c < V; c += blockDim.x) part[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r < rows) {
    const float* row = lm + r * V;
    const float m = lm_max[r];
    float z = 0.f;
    for (int c = lane; c < V; c += 32) z += expf(row[c] - m);
    z = warp_sum(z);
    if (lane == 0) lmN[r] = logf(z) + m;
    const float inv = 1.f / z;
    for (int c = lane; c < V; c += 32) atomicAdd(part + c, expf(row[c] - m) * inv);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < V; c += blockDim.x)
    if (part[c] != 0.f) atomicAdd(usum + c, part[c]);
}

__global__ void unigram_finish_kernel(const float* __restrict__ usum, int V, float inv_rows, float* __restrict__ u,
                                      float* __restrict__ logu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= V) return;
  const float v = usum[c] * inv_rows + FLT_MIN;
  u[c] = v;
  logu[c] = logf(v);
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) am_norm_kernel(const float* __restrict__ am,
                                                                      const float* __restrict__ am_max,
                                                                      const float* __restrict__ u, int64_t rows, int V,
                                                                      float* __restrict__ amN) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= rows) return;
  const float* row = am + r * V;
  const float m = am_max[r];
  float z = 0.f;
  for (int c = lane; c < V; c += 32) z += expf(row[c] - m) * u[c];
  z = warp_sum(z);
  if (lane == 0) amN[r] = logf(z) + m;
}

// in-place mix; thread per (b, s, t), t fastest
__global__ void smooth_mix_kernel(const float* __restrict__ lm, const int64_t* __restrict__ sym,
                                  const float* __restrict__ nrm, const float* __restrict__ lmN,
                                  const float* __restrict__ amN, const float* __restrict__ logu, int B, int S, int T,
                                  int V, int blank, float c0, float ls, float as, float* __restrict__ px,
                                  float* __restrict__ py) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * (S + 1) * T) return;
  const int t = (int)(i % T);
  const int64_t bs = i / T;
  const int s = (int)(bs % (S + 1)), b = (int)(bs / (S + 1));
  const float nv = nrm[i], ln = lmN[bs], an = amN[(int64_t)b * T + t];
  const float* lrow = lm + bs * V;
  {
    const float p = py[i], lb = __ldg(lrow + blank);
    py[i] = c0 * p + ls * (lb - ln) + as * (p - lb + nv + logu[blank] - an);
  }
  if (s < S) {
    float* q = px + ((int64_t)b * S + s) * (T + 1) + t;
    const float p = *q;
    if (p != kNegInf) {  // the column of the last frame stays -inf (fix_for_boundary)
      const int c = (int)sym[(int64_t)b * S + s];
      const float lsym = __ldg(lrow + c);
      *q = c0 * p + ls * (lsym - ln) + as * (p - lsym + nv + logu[c] - an);
    }
  }
}

// A[b,s] = coef_b sum_t (occ_px + occ_py); H[sym] += coef_b sum_t occ_px, H[blank] += coef_b sum_t occ_py
__global__ void __launch_bounds__(32 * kWarpsPerBlock) smooth_row_sums_kernel(
    const float* __restrict__ occ_px, const float* __restrict__ occ_py, const int64_t* __restrict__ sym,
    const float* __restrict__ coef, int B, int S, int T, int blank, float* __restrict__ A, float* __restrict__ H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (w >= (int64_t)B * (S + 1)) return;
  const int s = (int)(w % (S + 1)), b = (int)(w / (S + 1));
  float xs = 0.f, ys = 0.f;
  for (int t = lane; t < T; t += 32) ys += occ_py[w * T + t];
  if (s < S)
    for (int t = lane; t < T; t += 32) xs += occ_px[((int64_t)b * S + s) * (T + 1) + t];
  xs = warp_sum(xs);
  ys = warp_sum(ys);
  if (lane == 0) {
    const float cf = coef[b];
    A[w] = cf * (xs + ys);
    if (s < S && xs != 0.f) atomicAdd(H + sym[(int64_t)b * S + s], cf * xs);
    if (ys != 0.f) atomicAdd(H + blank, cf * ys);
  }
}

// Bt[b,t] = coef_b sum_s (occ_px + occ_py)
__global__ void smooth_col_sums_kernel(const float* __restrict__ occ_px, const float* __restrict__ occ_py,
                                       const float* __restrict__ coef, int B, int S, int T, float* __restrict__ Bt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T) return;
  const int t = (int)(i % T), b = (int)(i / T);
  float acc = 0.f;
  for (int s = 0; s <= S; ++s) acc += occ_py[((int64_t)b * (S + 1) + s) * T + t];
  for (int s = 0; s < S; ++s) acc += occ_px[((int64_t)b * S + s) * (T + 1) + t];
  Bt[i] = coef[b] * acc;
}

// d_am += -a_s Bt q,  q[c] = exp(am[c]) u[c] / Z;   dLdu[c] += -a_s Bt exp(am[c]) / Z
__global__ void __launch_bounds__(32 * kWarpsPerBlock) am_only_bwd_kernel(
    const float* __restrict__ am, const float* __restrict__ am_max, const float* __restrict__ u,
    const float* __restrict__ Bt, int64_t rows, int V, float as, float* __restrict__ d_am, float* __restrict__ dLdu) {
  extern __shared__ float part[];  // V
  for (int c = threadIdx.x; c < V; c += blockDim.x) part[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r < rows) {
    const float bt = Bt[r];
    if (bt != 0.f) {
      const float* row = am + r * V;
      const float m = am_max[r];
      float z = 0.f;
      for (int c = lane; c < V; c += 32) z += expf(row[c] - m) * u[c];
      z = warp_sum(z);
      const float k = -as * bt / z;
      for (int c = lane; c < V; c += 32) {
        const float e = expf(row[c] - m) * k;
        d_am[r * V + c] += e * u[c];
        atomicAdd(part + c, e);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < V; c += blockDim.x)
    if (part[c] != 0.f) atomicAdd(dLdu + c, part[c]);
}

__global__ void dldu_finish_kernel(const float* __restrict__ H, const float* __restrict__ u, int V, float as,
                                   float* __restrict__ dLdu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < V) dLdu[c] += as * H[c] / u[c];
}

// d_lm += -l_s A p + (1/N) p (dLdu - <dLdu, p>),  p = softmax(lm row)
__global__ void __launch_bounds__(32 * kWarpsPerBlock) lm_only_bwd_kernel(
    const float* __restrict__ lm, const float* __restrict__ lm_max, const float* __restrict__ lmN,
    const float* __restrict__ A, const float* __restrict__ dLdu, int64_t rows, int V, float ls, float inv_rows,
    float* __restrict__ d_lm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= rows) return;
  const float* row = lm + r * V;
  const float ln = lmN[r];
  float dot = 0.f;
  for (int c = lane; c < V; c += 32) dot += dLdu[c] * expf(row[c] - ln);
  dot = warp_sum(dot);
  const float a = -ls * A[r];
  for (int c = lane; c < V; c += 32) {
    const float p = expf(row[c] - ln);
    d_lm[r * V + c] += a * p + inv_rows * p * (dLdu[c] - dot);
  }
}

__global__ void scale_coef_kernel(const float* __restrict__ coef, int B, float k, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = coef[b] * k;
}

}  // namespace

size_t simple_smooth_workspace_bytes(int B, int T, int S, int V) {
  return (2 * ((size_t)B * (S + 1) + (size_t)B * T) + 5 * (size_t)V + 3 * (size_t)B + 64) * sizeof(float);
}

// px / py (plain) -> interpolated, in place
int simple_smooth_forward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                          const float* lm_max, const float* nrm, int B, int T, int S, int V, int blank, float ls,
                          float as, void* ws, float* px, float* py, cudaStream_t st) {
  SmoothWs w = carve(ws, B, T, S, V);
  const int64_t rows_lm = (int64_t)B * (S + 1), rows_am = (int64_t)B * T;
  ProfScope prof("simple_smooth_kernels", st, 4);
  cudaMemsetAsync(w.usum, 0, (size_t)V * sizeof(float), st);
  lm_stats_kernel<<<(unsigned)((rows_lm + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock,
                    (size_t)V * sizeof(float), st>>>(lm, lm_max, rows_lm, V, w.lmN, w.usum);
  unigram_finish_kernel<<<(V + 255) / 256, 256, 0, st>>>(w.usum, V, 1.f / (float)rows_lm, w.u, w.logu);
  am_norm_kernel<<<(unsigned)((rows_am + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
      am, am_max, w.u, rows_am, V, w.amN);
  const int64_t total = rows_lm * T;
  if (total > 0)
    smooth_mix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(lm, sym, nrm, w.lmN, w.amN, w.logu, B, S, T, V,
                                                                       blank, 1.f - ls - as, ls, as, px, py);
  return check_launch("simple_smooth_forward");
}

// coef (B) scaled by k into the workspace: slot 0 / 1 / 2
float* simple_smooth_scaled_coef(const float* coef, int B, int T, int S, int V, float k, int slot, void* ws,
                                 cudaStream_t st) {
  SmoothWs w = carve(ws, B, T, S, V);
  float* out = w.Bt + (size_t)B * T + (size_t)slot * B;
  scale_coef_kernel<<<(B + 127) / 128, 128, 0, st>>>(coef, B, k, out);
  return out;
}

// adds the terms that flow through lmN, amN and the unigram to d_am / d_lm (the workspace still holds u, lmN of the
// forward pass)
int simple_smooth_backward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                           const float* lm_max, const float* occ_px, const float* occ_py, const float* coef, int B,
                           int T, int S, int V, int blank, float ls, float as, void* ws, float* d_am, float* d_lm,
                           cudaStream_t st) {
  SmoothWs w = carve(ws, B, T, S, V);
  const int64_t rows_lm = (int64_t)B * (S + 1), rows_am = (int64_t)B * T;
  ProfScope prof("simple_smooth_kernels", st, 6);
  cudaMemsetAsync(w.dLdu, 0, 2 * (size_t)V * sizeof(float), st);  // dLdu and H are adjacent
  smooth_row_sums_kernel<<<(unsigned)((rows_lm + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
      occ_px, occ_py, sym, coef, B, S, T, blank, w.A, w.H);
  if (as != 0.f) {
    smooth_col_sums_kernel<<<(unsigned)((rows_am + 255) / 256), 256, 0, st>>>(occ_px, occ_py, coef, B, S, T, w.Bt);
    am_only_bwd_kernel<<<(unsigned)((rows_am + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock,
                         (size_t)V * sizeof(float), st>>>(am, am_max, w.u, w.Bt, rows_am, V, as, d_am, w.dLdu);
    dldu_finish_kernel<<<(V + 255) / 256, 256, 0, st>>>(w.H, w.u, V, as, w.dLdu);
  }
  lm_only_bwd_kernel<<<(unsigned)((rows_lm + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
      lm, lm_max, w.lmN, w.A, w.dLdu, rows_lm, V, ls, 1.f / (float)rows_lm, d_lm);
  return check_launch("simple_smooth_backward");
}

}  // namespace s2t
