// nn.Linear on the tensor cores for the joiner's D -> V projections
// (/root/reference/model/joiner/joiner.py:41-42, 148-149: _enc_proj / _pre_proj), fp32 output
// (the simple loss downstream is fp32).  The forward runs as 3xTF32 (big/small operand split,
// fp32-level accuracy): am / lm feed the simple-loss lattice whose occupation probabilities pick the
// prune ranges by an argmax, and any visible perturbation of am / lm (bf16: 2^-9, plain tf32: 2^-11)
// flips near-tie frames to a neighbouring window.  The backward contractions take bf16 operands.
//
//   forward   y = x W^T + b          K-major tf32: A = x copied on the fly, B = fp32 pack(W) (rows n, cols k)
//   backward  dx = dy W              K-major:  A = pack(dy) (rows m, cols n), B = pack(W^T) (rows k, cols n)
//             dW = dy^T x            MN-major: the SAME pack(dy) and pack(x), contraction over the rows m
//             db = column sums of dy (a by-product of packing dy)
// pack(x) is written once in the forward call and reused by the backward call (workspace).
#include <stdlib.h>
#include "../../include/s2t_b200.h"
#include "tc_gemm.cuh"

namespace s2t {
namespace {

using namespace tc;

// K-major contractions run as 2-CTA clusters: the B operand of a stage is loaded half by each CTA and multicast
constexpr int kPair = S2T_PAIR;  // the forward projection runs on CTA pairs (cta_group::2); dx measured better on single CTAs

struct LinDims {
  int Mt, Np, Kp;       // row tiles of x; N, K padded to multiples of 256
  size_t px, pw, pdy, pwt;  // bf16 pack(x), fp32 pack(W), bf16 pack(dy), bf16 pack(W^T)
};

LinDims lin_dims(int64_t M, int N, int K) {
  LinDims d;
  d.Mt = (int)((M + 127) / 128);
  d.Np = ((N + 255) / 256) * 256;
  d.Kp = ((K + 255) / 256) * 256;
  d.px = (size_t)d.Mt * (d.Kp / 64) * kBlockBytes;
  d.pw = 2 * (size_t)(d.Np / 128) * (d.Kp / 32) * kBlockBytes;  // big | small (sized for the tf32 split; the f16 split needs half)
  d.pdy = (size_t)d.Mt * (d.Np / 64) * kBlockBytes;
  d.pwt = (size_t)(d.Kp / 128) * (d.Np / 64) * kBlockBytes;
  return d;
}

}  // namespace
}  // namespace s2t

using namespace s2t;

extern "C" {

// workspace = [pack(x) | pack(W) | pack(dy) | pack(W^T)]; pack(x) and pack(W^T) must survive until backward
size_t s2t_linear_workspace_bytes(int64_t M, int N, int K) {
  LinDims d = lin_dims(M, N, K);
  return d.px + d.pw + d.pdy + d.pwt + 1024;
}

int s2t_linear_fwd(const void* x, int x_dtype, const float* W, const float* b, int64_t M, int N, int K, void* ws, float* y,
                   float* row_max, void* stream) {
  S2T_REQUIRE(x_dtype == S2T_F32 || x_dtype == S2T_BF16, "linear_fwd: x must be fp32 or bf16 (dtype code %d)", x_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  LinDims d = lin_dims(M, N, K);
  uint8_t* px = (uint8_t*)ws;
  uint8_t* pw = px + d.px;
  // pack(x) for dW in backward is a by-product of the forward producer; only K padding it never visits needs zeros
  const bool f16 = getenv("S2T_B200_LINEAR_TF32") == nullptr;  // default: 3xF16 split; the 3xTF32 path stays selectable
  const int kstep = f16 ? 64 : 32;
  if (((K + kstep - 1) / kstep) * kstep < d.Kp) cudaMemsetAsync(px, 0, d.px, st);
  uint8_t* pw_small = pw + d.pw / 2;
  uint8_t* pwt = pw + d.pw + d.pdy;
  tc::StoreRowMajorEpi ep{y, N, (int)M, N, false, b};
  ep.row_max = row_max;
  tc::MnDebug extra;
  extra.b_small = pw_small;
  if (f16) {
    // W is pre-scaled by 2^6 so that the lo halves of typical weights (|w| ~ 1e-2) stay normal f16 numbers; the
    // epilogue multiplies the accumulator by 2^-6 (exact).  |w| < 1023 and |x| < 65504 are representable.
    constexpr float kWScale = 64.f;
    const tc::PackJob jobs[3] = {
        {W, K, 1, N, K, d.Np / 128, d.Kp / 64, pw, 3, kWScale},
        {W, K, 1, N, K, d.Np / 128, d.Kp / 64, pw_small, 4, kWScale},
        // W^T for dx in backward: rows k, cols n -> element (k, n) = W[n * K + k]
        {W, 1, K, K, N, d.Kp / 128, d.Np / 64, pwt, 0},
    };
    if (int rc = tc::pack_jobs(jobs, 3, st, row_max, M, kNegInf)) return rc;  // + row_max = -inf for the epilogue's atomic max
    ep.scale = 1.f / kWScale;
    if (x_dtype == S2T_BF16) {
      // bf16 activations are exact in f16 (inside its range): no residual part of x, two MMAs per product instead of three
      if (getenv("S2T_B200_LINEAR_3X") == nullptr) {
        tc::RowSplitProducerF16T<__nv_bfloat16, false> a{(const __nv_bfloat16*)x, K, M, K, px, d.Mt};
        return tc::launch_gemm_stream<256, 3, false, 5, kPair>(a, pw, d.Np / 128, d.Mt, d.Np / 256, (K + 63) / 64, 1, ep, st,
                                                               "tc_linear_fwd_gemm_3xf16", extra);
      }
      tc::RowSplitProducerF16T<__nv_bfloat16> a{(const __nv_bfloat16*)x, K, M, K, px, d.Mt};
      return tc::launch_gemm_stream<256, 2, false, 3, kPair>(a, pw, d.Np / 128, d.Mt, d.Np / 256, (K + 63) / 64, 1, ep, st,
                                                             "tc_linear_fwd_gemm_3xf16", extra);
    }
    tc::RowSplitProducerF16 a{(const float*)x, K, M, K, px, d.Mt};
    return tc::launch_gemm_stream<256, 2, false, 3, kPair>(a, pw, d.Np / 128, d.Mt, d.Np / 256, (K + 63) / 64, 1, ep, st,
                                                           "tc_linear_fwd_gemm_3xf16", extra);
  }
  S2T_REQUIRE(x_dtype == S2T_F32, "linear_fwd: the 3xTF32 path (S2T_B200_LINEAR_TF32) takes fp32 activations");
  {
    const tc::PackJob jobs[3] = {
        {W, K, 1, N, K, d.Np / 128, d.Kp / 32, pw, 1},
        {W, K, 1, N, K, d.Np / 128, d.Kp / 32, pw_small, 2},
        {W, 1, K, K, N, d.Kp / 128, d.Np / 64, pwt, 0},
    };
    if (int rc = tc::pack_jobs(jobs, 3, st, row_max, M, kNegInf)) return rc;
  }
  tc::RowCopyProducerF32 a{(const float*)x, K, M, K, true, px, d.Mt};
  return tc::launch_gemm_stream<256, 2, false, 2, kPair>(a, pw, d.Np / 128, d.Mt, d.Np / 256, (K + 31) / 32, 1, ep, st,
                                                  "tc_linear_fwd_gemm_3xtf32", extra);
}

// dx (M,K), dW (N,K), db (N) are overwritten; the upstream gradient is dy (+ dy2 when not null)
int s2t_linear_bwd(const float* dy, const float* dy2, const float* W, int64_t M, int N, int K, void* ws, void* dx,
                   int dx_dtype, float* dW, float* db, void* stream) {
  S2T_REQUIRE(dx_dtype == S2T_F32 || dx_dtype == S2T_BF16, "linear_bwd: dx must be fp32 or bf16 (dtype code %d)", dx_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  LinDims d = lin_dims(M, N, K);
  uint8_t* px = (uint8_t*)ws;
  uint8_t* pdy = px + d.px + d.pw;
  uint8_t* pwt = pdy + d.pdy;
  // pack(dy) also yields db = column sums of dy
  // dW and db neighbours in a flat gradient bucket (weight, then bias): one memset node in front of everything
  const bool one_memset = db == dW + (size_t)N * K;
  if (one_memset) zero_async(dW, ((size_t)N * K + N) * sizeof(float), st);
  else cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st);
  if (int rc = tc::pack_rows_colsum(dy, dy2, N, (int)M, N, d.Mt, d.Np / 64, pdy, db, st)) return rc;
  ForkJoin fj(st);  // dx and dW only share the packed dy: the weight gradient runs on a side stream
  cudaStream_t s_dw = dx ? fj.side(0) : st;
  if (dx) {
    tc::BulkA a{pdy, d.Mt};
    if (dx_dtype == S2T_BF16) {
      tc::StoreRowMajorBf16Epi ep{(__nv_bfloat16*)dx, K, (int)M, K};
      if (int rc = tc::launch_gemm_stream<256, 4, false, 0>(a, pwt, d.Kp / 128, d.Mt, d.Kp / 256, (N + 63) / 64, 1, ep, st,
                                                         "tc_linear_dx_gemm"))
        return rc;
    } else {
      tc::StoreRowMajorEpi ep{(float*)dx, K, (int)M, K, false, nullptr};
      if (int rc = tc::launch_gemm_stream<256, 4, false, 0>(a, pwt, d.Kp / 128, d.Mt, d.Kp / 256, (N + 63) / 64, 1, ep, st,
                                                         "tc_linear_dx_gemm"))
        return rc;
    }
  }
  {
    if (!one_memset) cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), s_dw);
    const int k_steps = d.Mt * 2;
    const int tiles = (d.Np / 128) * (d.Kp / 256);
    int splits = device_info().sms / (tiles > 0 ? tiles : 1);
    if (splits < 1) splits = 1;
    tc::BulkA a{pdy, d.Mt};
    tc::StoreRowMajorEpi ep{dW, K, N, K, true, nullptr};
    if (int rc = tc::launch_gemm_stream<256, 4, true, 0>(a, px, d.Mt, d.Np / 128, d.Kp / 256, k_steps, splits, ep, s_dw,
                                                      "tc_linear_dW_gemm"))
      return rc;
  }
  fj.join();
  return check_launch("linear_bwd");
}

}  // extern "C"
