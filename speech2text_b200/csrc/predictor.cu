// Stateless predictor front end: the step right before the hot path.
//
// Replaces the embedding lookup + depthwise Conv1d(kernel = context_size, groups = E, no bias) of
// /root/reference/model/predictor/stateless_predictor.py:74-99 (forward) -- the n-gram-like context layer whose
// output feeds Linear(E, D) and then the joiner's _pre_proj:
//     h[b, u, e] = sum_{k < C} conv_w[e, k] * emb[ctx[b, u + k], e],      u = 0 .. U  (ctx = state | blank | tokens)
// One fused gather-multiply-add per output row instead of embedding -> transpose -> conv -> transpose (four
// (B, U + C, E)-sized tensors in the reference); the Linear that follows runs on the tensor-core GEMM of
// linear_tc.cu.  HBM-bound: E*4 bytes written per row, the embedding table (N x E) stays in L2.
//
// Backward: d_conv_w[e, k] = sum_{b,u} d_h[b,u,e] emb[ctx[b,u+k], e] (per-thread partial sums over the CTA's rows,
// one atomic per (e, k) and CTA) and d_emb[tok, e] += conv_w[e, k] d_h[b,u,e] (16-byte reductions: rows of the same
// token meet in the table).
#include "../../include/s2t_b200.h"
#include "common.cuh"

namespace s2t {
namespace {

constexpr int kMaxContext = 8;  // context sizes are compile-time (register arrays): the reference's configs use 5

__device__ __forceinline__ int clamp_token(int64_t t, int N) { return (int)(t < 0 ? 0 : (t >= N ? N - 1 : t)); }

// thread = four consecutive embedding channels of one output row; a CTA walks rows with a grid stride
template <bool kVec, int C>
__global__ void __launch_bounds__(256) predictor_embed_conv_fwd_kernel(const float* __restrict__ emb,
                                                                       const float* __restrict__ conv_w,
                                                                       const int64_t* __restrict__ ctx, int B, int L,
                                                                       int E, int N, float* __restrict__ h) {
  const int U1 = L - C + 1;  // output positions per utterance
  const int64_t rows = (int64_t)B * U1;
  const int per_row = (E + 3) / 4;
  const int rows_per_cta = blockDim.x / per_row > 0 ? blockDim.x / per_row : 1;
  const int sub = threadIdx.x / per_row, e = (threadIdx.x % per_row) * 4;
  if (sub >= rows_per_cta || e >= E) return;
  float w[C][4];
#pragma unroll
  for (int k = 0; k < C; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) w[k][q] = (e + q < E) ? __ldg(conv_w + (int64_t)(e + q) * C + k) : 0.f;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_cta + sub; row < rows; row += (int64_t)gridDim.x * rows_per_cta) {
    const int b = (int)(row / U1), u = (int)(row % U1);
    const int64_t* c = ctx + (int64_t)b * L + u;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float* src = emb + (int64_t)clamp_token(c[k], N) * E + e;
      float x[4];
      if (kVec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = (e + q < E) ? __ldg(src + q) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(w[k][q], x[q], acc[q]);
    }
    float* dst = h + row * E + e;
    if (kVec) {
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (e + q < E) dst[q] = acc[q];
    }
  }
}

template <bool kVec, int C>
__global__ void __launch_bounds__(256) predictor_embed_conv_bwd_kernel(const float* __restrict__ emb,
                                                                       const float* __restrict__ conv_w,
                                                                       const int64_t* __restrict__ ctx,
                                                                       const float* __restrict__ d_h, int B, int L,
                                                                       int E, int N, float* __restrict__ d_emb,
                                                                       float* __restrict__ d_conv_w) {
  const int U1 = L - C + 1;
  const int64_t rows = (int64_t)B * U1;
  const int per_row = (E + 3) / 4;
  const int rows_per_cta = blockDim.x / per_row > 0 ? blockDim.x / per_row : 1;
  const int sub = threadIdx.x / per_row, e = (threadIdx.x % per_row) * 4;
  if (sub >= rows_per_cta || e >= E) return;
  float w[C][4], dw[C][4];
#pragma unroll
  for (int k = 0; k < C; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w[k][q] = (e + q < E) ? __ldg(conv_w + (int64_t)(e + q) * C + k) : 0.f;
      dw[k][q] = 0.f;
    }
  for (int64_t row = (int64_t)blockIdx.x * rows_per_cta + sub; row < rows; row += (int64_t)gridDim.x * rows_per_cta) {
    const int b = (int)(row / U1), u = (int)(row % U1);
    const int64_t* c = ctx + (int64_t)b * L + u;
    float g[4];
    const float* gp = d_h + row * E + e;
    if (kVec) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gp));
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = (e + q < E) ? __ldg(gp + q) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int tok = clamp_token(c[k], N);
      const float* src = emb + (int64_t)tok * E + e;
      float x[4];
      if (kVec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = (e + q < E) ? __ldg(src + q) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) dw[k][q] = fmaf(g[q], x[q], dw[k][q]);
      float* dst = d_emb + (int64_t)tok * E + e;
      if (kVec) {
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(w[k][0] * g[0], w[k][1] * g[1], w[k][2] * g[2], w[k][3] * g[3]));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (e + q < E) atomicAdd(dst + q, w[k][q] * g[q]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < C; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (e + q < E && dw[k][q] != 0.f) atomicAdd(d_conv_w + (int64_t)(e + q) * C + k, dw[k][q]);
}

int predictor_grid(int64_t rows, int E, int threads) {
  const int per_row = (E + 3) / 4;
  const int rows_per_cta = threads / per_row > 0 ? threads / per_row : 1;
  const int64_t ctas = (rows + rows_per_cta - 1) / rows_per_cta;
  const int cap = device_info().sms * 8;
  return (int)(ctas < cap ? ctas : cap);
}

}  // namespace
}  // namespace s2t

using namespace s2t;

extern "C" {

int s2t_predictor_embed_conv_fwd(const float* emb, const float* conv_w, const int64_t* ctx, int B, int L, int C, int E,
                                 int N, float* h, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B > 0 && C >= 1 && C <= kMaxContext && L >= C && E > 0 && N > 0,
              "predictor_embed_conv: bad dims B=%d L=%d C=%d (<= %d) E=%d N=%d", B, L, C, kMaxContext, E, N);
  S2T_REQUIRE((E + 3) / 4 <= 256, "predictor_embed_conv: embedding dim %d > 1024", E);
  const int64_t rows = (int64_t)B * (L - C + 1);
  const bool vec = (E % 4 == 0) && (((uintptr_t)emb | (uintptr_t)h) % 16 == 0);
  ProfScope prof("predictor_embed_conv_fwd_kernel", st);
  const int grid = predictor_grid(rows, E, 256);
  switch (C) {
#define S2T_FWD(c)                                                                                           \
  case c:                                                                                                    \
    if (vec) predictor_embed_conv_fwd_kernel<true, c><<<grid, 256, 0, st>>>(emb, conv_w, ctx, B, L, E, N, h); \
    else predictor_embed_conv_fwd_kernel<false, c><<<grid, 256, 0, st>>>(emb, conv_w, ctx, B, L, E, N, h);    \
    break;
    S2T_FWD(1) S2T_FWD(2) S2T_FWD(3) S2T_FWD(4) S2T_FWD(5) S2T_FWD(6) S2T_FWD(7) S2T_FWD(8)
#undef S2T_FWD
  }
  return check_launch("predictor_embed_conv_fwd_kernel");
}

int s2t_predictor_embed_conv_bwd(const float* emb, const float* conv_w, const int64_t* ctx, const float* d_h, int B,
                                 int L, int C, int E, int N, float* d_emb, float* d_conv_w, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B > 0 && C >= 1 && C <= kMaxContext && L >= C && E > 0 && N > 0,
              "predictor_embed_conv: bad dims B=%d L=%d C=%d (<= %d) E=%d N=%d", B, L, C, kMaxContext, E, N);
  S2T_REQUIRE((E + 3) / 4 <= 256, "predictor_embed_conv: embedding dim %d > 1024", E);
  const int64_t rows = (int64_t)B * (L - C + 1);
  const bool vec = (E % 4 == 0) && (((uintptr_t)emb | (uintptr_t)d_h | (uintptr_t)d_emb) % 16 == 0);
  cudaMemsetAsync(d_emb, 0, (size_t)N * E * sizeof(float), st);
  cudaMemsetAsync(d_conv_w, 0, (size_t)E * C * sizeof(float), st);
  ProfScope prof("predictor_embed_conv_bwd_kernel", st);
  const int grid = predictor_grid(rows, E, 256);
  switch (C) {
#define S2T_BWD(c)                                                                                                          \
  case c:                                                                                                                   \
    if (vec) predictor_embed_conv_bwd_kernel<true, c><<<grid, 256, 0, st>>>(emb, conv_w, ctx, d_h, B, L, E, N, d_emb, d_conv_w); \
    else predictor_embed_conv_bwd_kernel<false, c><<<grid, 256, 0, st>>>(emb, conv_w, ctx, d_h, B, L, E, N, d_emb, d_conv_w);    \
    break;
    S2T_BWD(1) S2T_BWD(2) S2T_BWD(3) S2T_BWD(4) S2T_BWD(5) S2T_BWD(6) S2T_BWD(7) S2T_BWD(8)
#undef S2T_BWD
  }
  return check_launch("predictor_embed_conv_bwd_kernel");
}

}  // extern "C"
