// Loss on *materialised* joiner logits (B, T, R, V): streaming log-sum-exp over
// the vocabulary with the target/blank gather, and the matching gradient.
//
// Replaces, for callers that hand over a real logits tensor,
//   * get_rnnt_logprobs_pruned inside k2.rnnt_loss_pruned
//     (/root/reference/model/loss/pruned_rnnt_loss.py:39-48), R = prune_range,
//     rows addressed through ranges (B, T, R);
//   * the fused log-softmax of torchaudio's rnnt_loss
//     (/root/reference/model/loss/rnnt_loss.py:42-44), R = U+1,
//     ranges == nullptr (slot r is symbol position r).
// Both kernels are single-pass HBM streams over the logits: one warp per
// (b, t, r) row, 128-bit loads, fp32 accumulation.  Algorithmic bytes:
// forward V*e read per row; backward V*e read + V*e written per row.
#include "common.cuh"

namespace s2t {
namespace {

template <typename T>
struct Vec;  // 16-byte vector of T
template <>
struct Vec<float> { static constexpr int N = 4; };
template <>
struct Vec<__nv_bfloat16> { static constexpr int N = 8; };
template <>
struct Vec<__half> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[Vec<T>::N]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < Vec<T>::N; ++i) out[i] = to_float<T>(e[i]);
}

__device__ __forceinline__ void online_update(float& m, float& s, float x) {
  // streaming log-sum-exp: keep (max, sum of exp(x - max))
  if (x > m) {
    s = s * expf(m - x) + 1.f;
    m = x;
  } else {
    s += expf(x - m);
  }
}

__device__ __forceinline__ int row_symbol(const int64_t* sym, const int64_t* ranges, int64_t row,
                                          int b, int r, int S, int blank, int* s_out) {
  int s = ranges ? (int)ranges[row] : r;
  *s_out = s;
  return (s >= 0 && s < S) ? (int)sym[(int64_t)b * S + s] : blank;
}

template <typename T>
__global__ void lse_gather_kernel(const T* __restrict__ logits, const int64_t* __restrict__ sym,
                                  const int64_t* __restrict__ ranges, const int64_t* __restrict__ boundary,
                                  int64_t row0, int64_t rows, int TR, int R, int V, int S, int blank,
                                  float delay_penalty, float* __restrict__ lse, float* __restrict__ px,
                                  float* __restrict__ py) {
  const int64_t lrow = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (lrow >= rows) return;
  const int64_t row = row0 + lrow;  // global (b, t, r) row; `logits` holds rows [row0, row0 + rows)
  const int lane = threadIdx.x % 32;
  const T* p = logits + lrow * V;
  constexpr int N = Vec<T>::N;
  float m = kNegInf, s = 0.f;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  int c0 = 0;
  if (aligned) {
    const int nvec = V / N;
    for (int i = lane; i < nvec; i += 32) {
      float v[N];
      load_vec<T>(p + (int64_t)i * N, v);
      float vm = v[0];
#pragma unroll
      for (int j = 1; j < N; ++j) vm = fmaxf(vm, v[j]);
      if (vm > m) {
        s *= expf(m - vm);
        m = vm;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) s += expf(v[j] - m);
    }
    c0 = nvec * N;
  }
  for (int c = c0 + lane; c < V; c += 32) online_update(m, s, to_float<T>(p[c]));
  // merge the 32 (max, sum) pairs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    float mm = fmaxf(m, m2);
    float a = (m == kNegInf) ? 0.f : s * expf(m - mm);
    float b2 = (m2 == kNegInf) ? 0.f : s2 * expf(m2 - mm);
    s = a + b2;
    m = mm;
  }
  if (lane == 0) {
    const float l = m + logf(s);
    const int b = (int)(row / TR);
    const int t = (int)((row % TR) / R);
    const int r = (int)(row % R);
    int spos;
    const int c = row_symbol(sym, ranges, row, b, r, S, blank, &spos);
    lse[row] = l;
    float xv = to_float<T>(p[c]) - l;
    if (delay_penalty != 0.f) {
      const int Tb = boundary ? (int)boundary[4 * b + 3] : (TR / R);
      xv += delay_penalty * (0.5f * (float)(Tb - 1) - (float)t);
    }
    px[row] = xv;
    py[row] = to_float<T>(p[blank]) - l;
  }
}

// grad[row, c] = coef_b * (occ_px [c == sym] + occ_py [c == blank] - (occ_px + occ_py) * softmax_c)
// with coef_b = d loss / d scores[b]  (SURVEY.md A.7 with coef_b = -g/B for reduction='mean')
// `grad` may alias `logits` (element-wise in place).
template <typename T>
__global__ void logits_grad_kernel(const T* logits, const int64_t* __restrict__ sym,
                                   const int64_t* __restrict__ ranges, const float* __restrict__ lse,
                                   const float* __restrict__ occ_px, const float* __restrict__ occ_py,
                                   const float* __restrict__ coef, int64_t row0, int64_t rows, int TR, int R,
                                   int V, int S, int blank, float clamp, T* grad) {
  const int64_t lrow = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (lrow >= rows) return;
  const int64_t row = row0 + lrow;
  const int lane = threadIdx.x % 32;
  const int b = (int)(row / TR);
  const int r = (int)(row % R);
  int spos;
  const int csym = row_symbol(sym, ranges, row, b, r, S, blank, &spos);
  const float cf = coef[b];
  const float ox = occ_px[row], oy = occ_py[row];
  const float g = ox + oy;
  const float l = lse[row];
  const T* p = logits + lrow * V;
  T* q = grad + lrow * V;
  if (g == 0.f || cf == 0.f) {  // dead row (padding frame / outside the lattice): exact zeros
    for (int c = lane; c < V; c += 32) q[c] = from_float<T>(0.f);
    return;
  }
  for (int c = lane; c < V; c += 32) {
    float v = -g * expf(to_float<T>(p[c]) - l);  // d scores / d logits
    if (c == csym) v += ox;
    if (c == blank) v += oy;
    // torchaudio's `clamp` clips the per-utterance logit gradient before the upstream scale
    if (clamp > 0.f) v = fminf(fmaxf(v, -clamp), clamp);
    q[c] = from_float<T>(cf * v);
  }
}

}  // namespace

template <typename T>
int lse_gather_t(const void* logits, const int64_t* sym, const int64_t* ranges, const int64_t* boundary,
                 int B, int T_, int R, int V, int S, int blank, float delay_penalty, float* lse, float* px,
                 float* py, cudaStream_t stream) {
  int64_t rows = (int64_t)B * T_ * R;
  if (rows == 0) return 0;
  const int wpb = 8;
  ProfScope prof("lse_gather_kernel", stream);
  lse_gather_kernel<T><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      (const T*)logits, sym, ranges, boundary, 0, rows, T_ * R, R, V, S, blank, delay_penalty, lse, px, py);
  return check_launch("lse_gather_kernel");
}

template <typename T>
int logits_grad_t(const void* logits, const int64_t* sym, const int64_t* ranges, const float* lse,
                  const float* occ_px, const float* occ_py, const float* coef, int B, int T_, int R, int V,
                  int S, int blank, float clamp, void* grad, cudaStream_t stream) {
  int64_t rows = (int64_t)B * T_ * R;
  if (rows == 0) return 0;
  const int wpb = 8;
  ProfScope prof("logits_grad_kernel", stream);
  logits_grad_kernel<T><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      (const T*)logits, sym, ranges, lse, occ_px, occ_py, coef, 0, rows, T_ * R, R, V, S, blank, clamp, (T*)grad);
  return check_launch("logits_grad_kernel");
}

// row-chunk variants used by the strict-fp32 fused joiner (joiner_simt.cu)
int lse_gather_rows(const float* logits, const int64_t* sym, const int64_t* ranges, const int64_t* boundary,
                    int64_t row0, int64_t rows, int T, int R, int V, int S, int blank, float delay_penalty,
                    float* lse, float* px, float* py, cudaStream_t stream) {
  if (rows == 0) return 0;
  const int wpb = 8;
  ProfScope prof("lse_gather_kernel", stream);
  lse_gather_kernel<float><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      logits, sym, ranges, boundary, row0, rows, T * R, R, V, S, blank, delay_penalty, lse, px, py);
  return check_launch("lse_gather_kernel(rows)");
}

int logits_grad_rows(float* logits_inout, const int64_t* sym, const int64_t* ranges, const float* lse,
                     const float* occ_px, const float* occ_py, const float* coef, int64_t row0, int64_t rows,
                     int T, int R, int V, int S, int blank, float clamp, cudaStream_t stream) {
  if (rows == 0) return 0;
  const int wpb = 8;
  ProfScope prof("logits_grad_kernel", stream);
  logits_grad_kernel<float><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      logits_inout, sym, ranges, lse, occ_px, occ_py, coef, row0, rows, T * R, R, V, S, blank, clamp, logits_inout);
  return check_launch("logits_grad_kernel(rows)");
}

// dtype: 0 = fp32, 1 = bf16, 2 = fp16
int lse_gather(const void* logits, int dtype, const int64_t* sym, const int64_t* ranges,
               const int64_t* boundary, int B, int T, int R, int V, int S, int blank, float delay_penalty,
               float* lse, float* px, float* py, cudaStream_t stream) {
  switch (dtype) {
    case 0: return lse_gather_t<float>(logits, sym, ranges, boundary, B, T, R, V, S, blank, delay_penalty, lse, px, py, stream);
    case 1: return lse_gather_t<__nv_bfloat16>(logits, sym, ranges, boundary, B, T, R, V, S, blank, delay_penalty, lse, px, py, stream);
    case 2: return lse_gather_t<__half>(logits, sym, ranges, boundary, B, T, R, V, S, blank, delay_penalty, lse, px, py, stream);
  }
  set_error("lse_gather: unsupported dtype code %d", dtype);
  return 1;
}

int logits_grad(const void* logits, int dtype, const int64_t* sym, const int64_t* ranges, const float* lse,
                const float* occ_px, const float* occ_py, const float* coef, int B, int T, int R, int V, int S,
                int blank, float clamp, void* grad, cudaStream_t stream) {
  switch (dtype) {
    case 0: return logits_grad_t<float>(logits, sym, ranges, lse, occ_px, occ_py, coef, B, T, R, V, S, blank, clamp, grad, stream);
    case 1: return logits_grad_t<__nv_bfloat16>(logits, sym, ranges, lse, occ_px, occ_py, coef, B, T, R, V, S, blank, clamp, grad, stream);
    case 2: return logits_grad_t<__half>(logits, sym, ranges, lse, occ_px, occ_py, coef, B, T, R, V, S, blank, clamp, grad, stream);
  }
  set_error("logits_grad: unsupported dtype code %d", dtype);
  return 1;
}

}  // namespace s2t
