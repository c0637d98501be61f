// extern "C" surface of libs2t_b200.so (declared in include/s2t_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/s2t_b200.h"
#include "joiner.cuh"
#include "lattice.cuh"

namespace s2t {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

// implemented in the other translation units
int simple_logprobs(const float* am, const float* lm, const int64_t* sym, const int64_t* boundary, int B, int T,
                    int S, int V, int blank, float* am_max, float* lm_max, float* px, float* py, float* nrm,
                    cudaStream_t stream);
int simple_backward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                    const float* lm_max, const float* nrm, const float* occ_px, const float* occ_py,
                    const float* coef, int B, int T, int S, int V, int blank, float* wbuf, float* d_am,
                    float* d_lm, cudaStream_t stream);
size_t simple_tc_workspace_bytes(int B, int T, int S, int V);
int simple_logprobs_tc(const float* am, const float* lm, const int64_t* sym, const int64_t* boundary, int B, int T,
                       int S, int V, int blank, float* am_max, float* lm_max, float* px, float* py, float* nrm,
                       void* ws, int ready, cudaStream_t stream);
int simple_prep_lm_tc(const float* lm, const float* lm_max, const int64_t* sym, int B, int T, int S, int V, int blank,
                      void* ws, cudaStream_t stream);
int simple_backward_tc(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                       const float* lm_max, const float* nrm, const float* occ_px, const float* occ_py,
                       const float* coef, int B, int T, int S, int V, int blank, void* ws, float* d_am, float* d_lm,
                       cudaStream_t stream);
size_t simple_smooth_workspace_bytes(int B, int T, int S, int V);
int simple_smooth_forward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                          const float* lm_max, const float* nrm, int B, int T, int S, int V, int blank, float ls,
                          float as, void* ws, float* px, float* py, cudaStream_t st);
float* simple_smooth_scaled_coef(const float* coef, int B, int T, int S, int V, float k, int slot, void* ws,
                                 cudaStream_t st);
int simple_smooth_backward(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                           const float* lm_max, const float* occ_px, const float* occ_py, const float* coef, int B,
                           int T, int S, int V, int blank, float ls, float as, void* ws, float* d_am, float* d_lm,
                           cudaStream_t st);
int simple_scatter_onehot_split(const float* occ_px, const float* occ_py, const int64_t* sym, const float* coef_am,
                                const float* coef_lm, int B, int S, int T, int V, int blank, float* d_am, float* d_lm,
                                cudaStream_t stream);
int prune_ranges(const float* px_grad, const float* py_grad, const int64_t* boundary, int B, int S, int T, int R,
                 int variant, int64_t* ranges, cudaStream_t stream);
int lse_gather(const void* logits, int dtype, const int64_t* sym, const int64_t* ranges, const int64_t* boundary,
               int B, int T, int R, int V, int S, int blank, float delay_penalty, float* lse, float* px, float* py,
               cudaStream_t stream);
int logits_grad(const void* logits, int dtype, const int64_t* sym, const int64_t* ranges, const float* lse,
                const float* occ_px, const float* occ_py, const float* coef, int B, int T, int R, int V, int S,
                int blank, float clamp, void* grad, cudaStream_t stream);

static LatticeView simple_view(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                               void* ws) {
  LatticeView v{};
  v.px = px;
  v.py = py;
  v.px_bs = (int64_t)S * (T + 1);
  v.px_ts = 1;
  v.px_rs = T + 1;
  v.py_bs = (int64_t)(S + 1) * T;
  v.py_ts = 1;
  v.py_rs = T;
  v.rx = S;
  v.ry = S + 1;
  v.px_at_tb = true;
  v.ranges = nullptr;
  v.boundary = boundary;
  v.B = B;
  v.S = S;
  v.T = T;
  lattice_carve_workspace(v, ws, S + 1);
  v.a_bs = (int64_t)(S + 1) * (T + 1);
  v.a_ts = 1;
  v.a_rs = T + 1;
  return v;
}

static LatticeView band_view(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary,
                             int B, int S, int T, int R, void* ws) {
  LatticeView v{};
  v.px = px;
  v.py = py;
  v.px_bs = v.py_bs = (int64_t)T * R;
  v.px_ts = v.py_ts = R;
  v.px_rs = v.py_rs = 1;
  v.rx = v.ry = R;
  v.px_at_tb = false;
  v.ranges = ranges;
  v.rg_bs = (int64_t)T * R;
  v.rg_ts = R;
  v.boundary = boundary;
  v.B = B;
  v.S = S;
  v.T = T;
  lattice_carve_workspace(v, ws, R);
  v.a_bs = (int64_t)(T + 1) * R;
  v.a_ts = R;
  v.a_rs = 1;
  return v;
}

static int band_dp(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary, int B, int S,
                   int T, int R, void* alpha, float* scores, float* occ_px, float* occ_py, cudaStream_t st) {
  if (band_lattice_fast_ok(S, T, R) && !getenv("S2T_B200_GENERIC_DP")) {
    return launch_band_lattice_fast(px, py, ranges, boundary, B, S, T, R, scores, occ_px, occ_py, st);
  }
  if (ranges == nullptr && R == S + 1 && full_lattice_fast_ok(S, T) && !getenv("S2T_B200_GENERIC_DP")) {
    return launch_full_lattice_fast(px, py, boundary, B, S, T, alpha, scores, occ_px, occ_py, st);
  }
  LatticeView v = band_view(px, py, ranges, boundary, B, S, T, R, alpha);
  size_t n = (size_t)B * T * R * sizeof(float);
  cudaMemsetAsync(occ_px, 0, n, st);
  cudaMemsetAsync(occ_py, 0, n, st);
  return launch_lattice_fwd_bwd(v, scores, occ_px, occ_py, st);
}

static JoinerProblem make_problem(const float* am, const float* lm, const int64_t* symbols, const int64_t* ranges,
                                  const int64_t* boundary, const float* W1, const float* b1, const float* W2,
                                  const float* b2, int B, int T, int S, int R, int V, int I, int act, int blank,
                                  float delay_penalty) {
  JoinerProblem p{};
  p.am = am; p.lm = lm; p.sym = symbols; p.ranges = ranges; p.boundary = boundary;
  p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2;
  p.B = B; p.T = T; p.S = S; p.R = R; p.V = V; p.I = I;
  p.act = act; p.blank = blank; p.delay_penalty = delay_penalty;
  return p;
}

}  // namespace s2t

namespace s2t {
namespace {

// x[g, :] *= num[g] / den[g] where the two differ; a block whose group needs no correction leaves after two loads
__global__ void __launch_bounds__(256) rescale_groups_kernel(float* __restrict__ x0, int64_t n0, float* __restrict__ x1,
                                                             int64_t n1, const float* __restrict__ num,
                                                             const float* __restrict__ den) {
  const int g = blockIdx.y;
  const float a = num[g], b = den[g];
  if (a == b) return;
  const float r = a / b;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (x0) {
    float* p = x0 + (int64_t)g * n0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += stride) p[i] *= r;
  }
  if (x1) {
    float* p = x1 + (int64_t)g * n1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) p[i] *= r;
  }
}

__global__ void __launch_bounds__(1024) weighted_sum_kernel(const float* __restrict__ x, const float* __restrict__ w, int n,
                                                            float* __restrict__ out) {
  __shared__ float red[32];
  const int i = threadIdx.x;
  float v = i < n ? x[i] * w[i] : 0.f;
  v = warp_sum(v);
  if ((i & 31) == 0) red[i >> 5] = v;
  __syncthreads();
  if (i < 32) {
    v = i < (int)((blockDim.x + 31) >> 5) ? red[i] : 0.f;
    v = warp_sum(v);
    if (i == 0) out[0] = v;
  }
}

__global__ void rescale_update_kernel(const float* __restrict__ num, float* __restrict__ den, int groups) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < groups) den[g] = num[g] != 0.f ? num[g] : 1.f;
}

}  // namespace
}  // namespace s2t

using namespace s2t;

extern "C" {

int s2t_abi_version(void) { return S2T_ABI_VERSION; }
const char* s2t_last_error(void) { return g_error; }

int s2t_mutual_information(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                           void* alpha_ws, float* scores, float* px_grad, float* py_grad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B >= 0 && S >= 0 && T >= 0, "mutual_information: negative dimension");
  if (simple_lattice_fast_ok(S, T) && !getenv("S2T_B200_GENERIC_DP")) {
    const bool grads = px_grad != nullptr && py_grad != nullptr;
    return launch_simple_lattice_fast(px, py, boundary, B, S, T, alpha_ws, scores, grads ? px_grad : nullptr,
                                      grads ? py_grad : nullptr, st);
  }
  LatticeView v = simple_view(px, py, boundary, B, S, T, alpha_ws);
  if (px_grad == nullptr || py_grad == nullptr) return launch_lattice_fwd(v, scores, st);
  cudaMemsetAsync(px_grad, 0, (size_t)B * S * (T + 1) * sizeof(float), st);
  cudaMemsetAsync(py_grad, 0, (size_t)B * (S + 1) * T * sizeof(float), st);
  return launch_lattice_fwd_bwd(v, scores, px_grad, py_grad, st);
}

static size_t simple_base_workspace_bytes(int mode, int B, int T, int S, int V) {
  size_t n = mode == S2T_MODE_BF16_TC ? simple_tc_workspace_bytes(B, T, S, V)
                                      : (size_t)B * (S + 1) * T * sizeof(float) + 256;  // W scratch of the fp32 backward
  return (n + 255) / 256 * 256;
}

size_t s2t_simple_workspace_bytes(int mode, int B, int T, int S, int V) {
  return simple_base_workspace_bytes(mode, B, T, S, V) + simple_smooth_workspace_bytes(B, T, S, V);
}

int s2t_simple_loss_fwd(int mode, const float* am, const float* lm, const int64_t* symbols, const int64_t* boundary,
                        int B, int T, int S, int V, int blank, float lm_only_scale, float am_only_scale,
                        float* am_max, float* lm_max, float* px, float* py, float* nrm, void* alpha_ws,
                        float* scores, float* px_grad, float* py_grad, void* workspace, int row_max_ready, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(lm_only_scale >= 0.f && am_only_scale >= 0.f && lm_only_scale + am_only_scale < 1.f,
              "simple_loss: lm_only_scale/am_only_scale (%g, %g) must be >= 0 and sum to less than 1", lm_only_scale,
              am_only_scale);
  S2T_REQUIRE(B > 0 && T > 0 && S >= 0 && V > 0, "simple_loss: bad dims B=%d T=%d S=%d V=%d", B, T, S, V);
  S2T_REQUIRE(blank >= 0 && blank < V, "simple_loss: blank %d out of range", blank);
  if (mode == S2T_MODE_BF16_TC) {
    if (int rc = simple_logprobs_tc(am, lm, symbols, boundary, B, T, S, V, blank, am_max, lm_max, px, py, nrm,
                                    workspace, row_max_ready, st))
      return rc;
  } else {
    // the fp32 path forms its own row maxima (it overwrites am_max / lm_max)
    if (int rc = simple_logprobs(am, lm, symbols, boundary, B, T, S, V, blank, am_max, lm_max, px, py, nrm, st))
      return rc;
  }
  if (lm_only_scale != 0.f || am_only_scale != 0.f) {
    void* sws = (char*)workspace + simple_base_workspace_bytes(mode, B, T, S, V);
    if (int rc = simple_smooth_forward(am, lm, symbols, am_max, lm_max, nrm, B, T, S, V, blank, lm_only_scale,
                                       am_only_scale, sws, px, py, st))
      return rc;
  }
  return s2t_mutual_information(px, py, boundary, B, S, T, alpha_ws, scores, px_grad, py_grad, stream);
}

int s2t_simple_loss_prep_lm(int mode, const float* lm, const float* lm_max, const int64_t* symbols, int B, int T, int S,
                            int V, int blank, void* workspace, void* stream) {
  S2T_REQUIRE(mode == S2T_MODE_BF16_TC, "simple_loss_prep_lm: tensor-core mode only (mode %d)", mode);
  S2T_REQUIRE(B > 0 && T > 0 && S >= 0 && V > 0 && blank >= 0 && blank < V, "simple_loss_prep_lm: bad dims B=%d T=%d S=%d V=%d", B,
              T, S, V);
  return simple_prep_lm_tc(lm, lm_max, symbols, B, T, S, V, blank, workspace, (cudaStream_t)stream);
}

int s2t_simple_loss_bwd(int mode, const float* am, const float* lm, const int64_t* symbols, const float* am_max,
                        const float* lm_max, const float* nrm, const float* px_grad, const float* py_grad,
                        const float* grad_scores, int B, int T, int S, int V, int blank, float lm_only_scale,
                        float am_only_scale, void* workspace, float* d_am, float* d_lm, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool smooth = lm_only_scale != 0.f || am_only_scale != 0.f;
  void* sws = (char*)workspace + simple_base_workspace_bytes(mode, B, T, S, V);
  // px_i holds the plain term with weight c0 and the am / lm one-hots with c0 + am_only_scale / c0 + lm_only_scale
  const float* coef = grad_scores;
  if (smooth)
    coef = simple_smooth_scaled_coef(grad_scores, B, T, S, V, 1.f - lm_only_scale - am_only_scale, 0, sws, st);
  int rc;
  if (mode == S2T_MODE_BF16_TC) {
    rc = simple_backward_tc(am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, coef, B, T, S, V, blank, workspace,
                            d_am, d_lm, st);
  } else {
    rc = simple_backward(am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, coef, B, T, S, V, blank,
                         (float*)workspace, d_am, d_lm, st);
  }
  if (rc || !smooth) return rc;
  const float* ca = am_only_scale != 0.f ? simple_smooth_scaled_coef(grad_scores, B, T, S, V, am_only_scale, 1, sws, st)
                                         : nullptr;
  const float* cl = lm_only_scale != 0.f ? simple_smooth_scaled_coef(grad_scores, B, T, S, V, lm_only_scale, 2, sws, st)
                                         : nullptr;
  if (int rc2 = simple_scatter_onehot_split(px_grad, py_grad, symbols, ca, cl, B, S, T, V, blank, d_am, d_lm, st))
    return rc2;
  return simple_smooth_backward(am, lm, symbols, am_max, lm_max, px_grad, py_grad, grad_scores, B, T, S, V, blank,
                                lm_only_scale, am_only_scale, sws, d_am, d_lm, st);
}

int s2t_prune_ranges(const float* px_grad, const float* py_grad, const int64_t* boundary, int B, int S, int T,
                     int s_range, int variant, int64_t* ranges, void* stream) {
  S2T_REQUIRE(boundary != nullptr, "prune_ranges: boundary is required");
  return prune_ranges(px_grad, py_grad, boundary, B, S, T, s_range, variant, ranges, (cudaStream_t)stream);
}

int s2t_logits_loss_fwd(const void* logits, int dtype, const int64_t* symbols, const int64_t* ranges,
                        const int64_t* boundary, int B, int T, int S, int R, int V, int blank,
                        float delay_penalty, float* lse, float* px, float* py, void* alpha_ws, float* scores,
                        float* occ_px, float* occ_py, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(ranges != nullptr || R == S + 1, "logits_loss: unpruned logits need R == S+1 (R=%d, S=%d)", R, S);
  if (int rc = lse_gather(logits, dtype, symbols, ranges, boundary, B, T, R, V, S, blank, delay_penalty, lse, px,
                          py, st))
    return rc;
  return band_dp(px, py, ranges, boundary, B, S, T, R, alpha_ws, scores, occ_px, occ_py, st);
}

int s2t_logits_loss_bwd(const void* logits, int dtype, const int64_t* symbols, const int64_t* ranges,
                        const float* lse, const float* occ_px, const float* occ_py, const float* grad_scores,
                        int B, int T, int S, int R, int V, int blank, float clamp, void* grad, void* stream) {
  return logits_grad(logits, dtype, symbols, ranges, lse, occ_px, occ_py, grad_scores, B, T, R, V, S, blank, clamp,
                     grad, (cudaStream_t)stream);
}

size_t s2t_lattice_workspace_bytes(int B, int S, int T, int slots) {
  size_t generic = lattice_workspace_bytes(B, S, T, slots);
  size_t fast = simple_lattice_fast_workspace_bytes(B, S, T);
  size_t full = slots == S + 1 ? full_lattice_fast_workspace_bytes(B, S, T) : 0;
  size_t best = generic > fast ? generic : fast;
  return best > full ? best : full;
}

// use_out_project=False (the reference's zipformer yaml): the joiner has no contraction, its "logits" are
// act(am + lm[ranges]) -- an HBM-bound elementwise + log-sum-exp pass that is the same code in both modes (the
// tensor-core mode then still covers the projections and the simple loss)
static bool joiner_uses_tc(int mode, int I) { return mode == S2T_MODE_BF16_TC && I > 0; }

size_t s2t_joiner_workspace_bytes(int mode, int B, int T, int R, int V, int I) {
  if (joiner_uses_tc(mode, I)) return joiner_tc_workspace_bytes((int64_t)B * T * R, V, I);
  return joiner_simt_workspace_bytes((int64_t)B * T * R, V, I, nullptr);
}

int s2t_weighted_sum(const float* x, const float* w, int n, float* out, void* stream) {
  S2T_REQUIRE(n >= 1 && n <= 1024 && x && w && out, "weighted_sum: n = %d (1..1024)", n);
  const int threads = ((n + 31) / 32) * 32;
  weighted_sum_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(x, w, n, out);
  return check_launch("weighted_sum_kernel");
}

int s2t_rescale_groups(float* x0, int64_t n0, float* x1, int64_t n1, int groups, const float* num, float* den,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(groups > 0 && groups <= 65535 && num && den, "rescale_groups: bad arguments (groups=%d)", groups);
  {
    ProfScope prof("rescale_groups_kernel", st);
    const int64_t n = n0 > n1 ? n0 : n1;
    int bx = (int)((n + 256 * 8 - 1) / (256 * 8));
    const int cap = device_info().sms * 8 / groups + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    rescale_groups_kernel<<<dim3((unsigned)bx, (unsigned)groups), 256, 0, st>>>(x0, n0, x1, n1, num, den);
    rescale_update_kernel<<<(groups + 255) / 256, 256, 0, st>>>(num, den, groups);
  }
  return check_launch("rescale_groups_kernel");
}

int s2t_joiner_logprobs_fwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                            const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                            const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                            int blank, float delay_penalty, void* workspace, float* lse, float* px, float* py,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(mode == S2T_MODE_FP32_SIMT || mode == S2T_MODE_BF16_TC, "joiner_loss_fwd: unknown mode %d", mode);
  S2T_REQUIRE(ranges != nullptr || R == S + 1, "joiner_loss: unpruned joiner needs R == S+1 (R=%d, S=%d)", R, S);
  S2T_REQUIRE(I == 0 || (W1 && b1 && W2 && b2), "joiner_loss: out-projection weights missing");
  JoinerProblem p = make_problem(am, lm, symbols, ranges, boundary, W1, b1, W2, b2, B, T, S, R, V, I, act, blank,
                                 delay_penalty);
  if (joiner_uses_tc(mode, I)) return joiner_tc_forward(p, workspace, lse, px, py, st);
  return joiner_simt_forward(p, workspace, lse, px, py, st);
}

int s2t_band_lattice_fwd(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary, int B,
                         int S, int T, int R, void* alpha_ws, float* scores, float* occ_px, float* occ_py,
                         void* stream) {
  S2T_REQUIRE(ranges != nullptr || R == S + 1, "band_lattice: unpruned lattice needs R == S+1 (R=%d, S=%d)", R, S);
  return band_dp(px, py, ranges, boundary, B, S, T, R, alpha_ws, scores, occ_px, occ_py, (cudaStream_t)stream);
}

int s2t_joiner_loss_fwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                        const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                        const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                        int blank, float delay_penalty, void* workspace, float* lse, float* px, float* py,
                        void* alpha_ws, float* scores, float* occ_px, float* occ_py, void* stream) {
  if (int rc = s2t_joiner_logprobs_fwd(mode, am, lm, symbols, ranges, boundary, W1, b1, W2, b2, B, T, S, R, V, I, act, blank,
                                       delay_penalty, workspace, lse, px, py, stream))
    return rc;
  return s2t_band_lattice_fwd(px, py, ranges, boundary, B, S, T, R, alpha_ws, scores, occ_px, occ_py, stream);
}

int s2t_joiner_loss_bwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                        const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                        const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                        int blank, float clamp, void* workspace, const float* lse, const float* occ_px,
                        const float* occ_py, const float* grad_scores, float* d_am, float* d_lm, float* dW1,
                        float* db1, float* dW2, float* db2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(mode == S2T_MODE_FP32_SIMT || mode == S2T_MODE_BF16_TC, "joiner_loss_bwd: unknown mode %d", mode);
  JoinerProblem p = make_problem(am, lm, symbols, ranges, boundary, W1, b1, W2, b2, B, T, S, R, V, I, act, blank,
                                 0.f);
  const bool tc = joiner_uses_tc(mode, I);
  if (!tc) cudaMemsetAsync(d_am, 0, (size_t)B * T * V * sizeof(float), st);  // the tensor-core path writes d_am itself
  zero_async(d_lm, (size_t)B * (S + 1) * V * sizeof(float), st);
  if (I > 0) {
    if (db1 == dW1 + (size_t)I * V && dW2 == db1 + I && db2 == dW2 + (size_t)V * I) {
      // the four gradients are neighbours in a flat gradient bucket (FlatGradBucket.bind): one memset node
      zero_async(dW1, ((size_t)2 * I * V + I + V) * sizeof(float), st);
    } else {
      cudaMemsetAsync(dW1, 0, (size_t)I * V * sizeof(float), st);
      cudaMemsetAsync(db1, 0, (size_t)I * sizeof(float), st);
      cudaMemsetAsync(dW2, 0, (size_t)V * I * sizeof(float), st);
      cudaMemsetAsync(db2, 0, (size_t)V * sizeof(float), st);
    }
  }
  if (joiner_uses_tc(mode, I)) {
    return joiner_tc_backward(p, workspace, lse, occ_px, occ_py, grad_scores, clamp, d_am, d_lm, dW1, db1, dW2, db2,
                              st);
  }
  return joiner_simt_backward(p, workspace, lse, occ_px, occ_py, grad_scores, clamp, d_am, d_lm, dW1, db1, dW2, db2,
                              st);
}

int s2t_joiner_materialize(int mode, const float* am, const float* lm, const int64_t* ranges, const float* W1,
                           const float* b1, const float* W2, const float* b2, int B, int T, int S, int R, int V,
                           int I, int act, void* workspace, float* logits, void* stream) {
  S2T_REQUIRE(mode == S2T_MODE_FP32_SIMT, "joiner_materialize: mode %d not built", mode);
  JoinerProblem p = make_problem(am, lm, nullptr, ranges, nullptr, W1, b1, W2, b2, B, T, S, R, V, I, act, 0, 0.f);
  return joiner_simt_materialize(p, workspace, logits, (cudaStream_t)stream);
}

}  // extern "C"
