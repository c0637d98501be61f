// Strict-fp32 pruned joiner + loss front end (SIMT mode).
//
// Replaces, without ever holding more than one row-chunk of logits,
//   k2.do_rnnt_pruning                      /root/reference/model/joiner/joiner.py:121-123
//   add + ReLU/Tanh + Linear(V,I) + Linear(I,V)   joiner.py:176-178, 51-57
//   logsumexp + sym/blank gather of k2.rnnt_loss_pruned   model/loss/pruned_rnnt_loss.py:39-48 (A.6)
// and their gradients (A.7).  Rows are m = (b, t, r); the A operand
// act(am[b,t,:] + lm[b,ranges[b,t,r],:]) is built on the fly inside the
// contraction (sgemm.cuh), so am_pruned / lm_pruned / joint / activation tensors
// never exist.  This is the parity mode (fp32 FMA, 1e-5 / 1e-4 against the
// oracle); the tensor-core mode lives in joiner_tc.cu.
#include "joiner.cuh"
#include "sgemm.cuh"

namespace s2t {

int lse_gather_rows(const float* logits, const int64_t* sym, const int64_t* ranges, const int64_t* boundary,
                    int64_t row0, int64_t rows, int T, int R, int V, int S, int blank, float delay_penalty,
                    float* lse, float* px, float* py, cudaStream_t stream);
int logits_grad_rows(float* logits_inout, const int64_t* sym, const int64_t* ranges, const float* lse,
                     const float* occ_px, const float* occ_py, const float* coef, int64_t row0, int64_t rows,
                     int T, int R, int V, int S, int blank, float clamp, cudaStream_t stream);

namespace {

// per-row element offsets into am (B,T,V) and lm (B,S+1,V)
__global__ void row_offsets_kernel(const int64_t* __restrict__ ranges, int64_t rows, int T, int R, int S,
                                   int V, int64_t* __restrict__ am_off, int64_t* __restrict__ lm_off) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  int64_t bt = m / R;
  int r = (int)(m % R);
  int64_t b = bt / T;
  int s = ranges ? (int)ranges[m] : r;
  s = min(max(s, 0), S);
  am_off[m] = bt * V;
  lm_off[m] = (b * (S + 1) + s) * V;
}

// element(m, k = c) = act(am[am_off[m] + c] + lm[lm_off[m] + c])      (k contiguous)
struct JointRowOperand {
  const float* am;
  const float* lm;
  const int64_t* am_off;
  const int64_t* lm_off;
  int64_t row0;
  int act;
  struct Row { const float* a; const float* l; };
  __device__ Row row(int, int m) const { return Row{am + am_off[row0 + m], lm + lm_off[row0 + m]}; }
  __device__ float at(const Row& r, int k) const { return act_fwd(__ldg(r.a + k) + __ldg(r.l + k), act); }
};

// element(n = c, k = m) = act(am[am_off[m] + c] + lm[lm_off[m] + c])  (n contiguous)
struct JointColOperand {
  const float* am;
  const float* lm;
  const int64_t* am_off;
  const int64_t* lm_off;
  int64_t row0;
  int act;
  struct Row { int c; };
  __device__ Row row(int, int c) const { return Row{c}; }
  __device__ float at(const Row& r, int k) const {
    return act_fwd(__ldg(am + am_off[row0 + k] + r.c) + __ldg(lm + lm_off[row0 + k] + r.c), act);
  }
};

struct BiasStoreEpilogue {  // out[(row0 + m) * ld + n] = acc + bias[n]
  float* out;
  const float* bias;
  int64_t row0;
  int ld;
  __device__ void operator()(int, int m, int n, float acc) const {
    out[(row0 + m) * ld + n] = acc + (bias ? bias[n] : 0.f);
  }
};

struct AtomicAccEpilogue {  // out[m * ld + n] += acc   (split-K partials)
  float* out;
  int ld;
  __device__ void operator()(int, int m, int n, float acc) const { atomicAdd(out + (int64_t)m * ld + n, acc); }
};

struct JointGradEpilogue {  // dj = acc * act'(am + lm); scatter-add into d_am / d_lm
  const float* am;
  const float* lm;
  const int64_t* am_off;
  const int64_t* lm_off;
  int64_t row0;
  int act;
  float* d_am;
  float* d_lm;
  __device__ void operator()(int, int m, int c, float acc) const {
    int64_t ao = am_off[row0 + m] + c, lo = lm_off[row0 + m] + c;
    float dj = acc * act_bwd(__ldg(am + ao) + __ldg(lm + lo), act);
    if (dj != 0.f) {
      atomicAdd(d_am + ao, dj);
      atomicAdd(d_lm + lo, dj);
    }
  }
};

// logits[m, c] = act(am + lm) for the no-out-projection joiner
__global__ void joint_act_kernel(const float* __restrict__ am, const float* __restrict__ lm,
                                 const int64_t* __restrict__ am_off, const int64_t* __restrict__ lm_off,
                                 int64_t row0, int64_t rows, int V, int act, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * V) return;
  int64_t m = i / V;
  int c = (int)(i % V);
  out[i] = act_fwd(__ldg(am + am_off[row0 + m] + c) + __ldg(lm + lm_off[row0 + m] + c), act);
}

// no-out-projection backward: dj = g[m, c] * act'(am + lm) -> d_am, d_lm
__global__ void joint_grad_kernel(const float* __restrict__ g, const float* __restrict__ am,
                                  const float* __restrict__ lm, const int64_t* __restrict__ am_off,
                                  const int64_t* __restrict__ lm_off, int64_t row0, int64_t rows, int V, int act,
                                  float* __restrict__ d_am, float* __restrict__ d_lm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * V) return;
  int64_t m = i / V;
  int c = (int)(i % V);
  float gv = g[i];
  if (gv == 0.f) return;
  int64_t ao = am_off[row0 + m] + c, lo = lm_off[row0 + m] + c;
  float dj = gv * act_bwd(__ldg(am + ao) + __ldg(lm + lo), act);
  if (dj != 0.f) {
    atomicAdd(d_am + ao, dj);
    atomicAdd(d_lm + lo, dj);
  }
}

// out[n] += sum_m x[m * ld + n]
__global__ void col_sum_kernel(const float* __restrict__ x, int64_t rows, int ld, int N, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t r0 = (int64_t)blockIdx.y * 256, r1 = min(rows, r0 + 256);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += x[r * ld + n];
  atomicAdd(out + n, acc);
}

int64_t pick_chunk_rows(int64_t M, int V, size_t logits_bytes) {
  int64_t rows = (int64_t)(logits_bytes / ((size_t)V * sizeof(float)));
  rows = (rows / 128) * 128;
  if (rows < 128) rows = 128;
  return rows < M ? rows : M;
}

}  // namespace

// Workspace layout (floats unless noted):
//   am_off, lm_off : M int64 each
//   hidden         : M * I            (kept from forward to backward)
//   logits chunk   : chunk_rows * V
//   dhid chunk     : chunk_rows * I
size_t joiner_simt_workspace_bytes(int64_t M, int V, int I, int64_t* chunk_rows_out) {
  const size_t budget = (size_t)1 << 30;  // 1 GiB of chunk logits at most
  int64_t chunk = pick_chunk_rows(M, V, budget);
  if (chunk_rows_out) *chunk_rows_out = chunk;
  size_t bytes = 0;
  bytes += 2 * (size_t)M * sizeof(int64_t);
  bytes += (size_t)M * (I > 0 ? I : 0) * sizeof(float);
  bytes += (size_t)chunk * V * sizeof(float);
  bytes += (size_t)chunk * (I > 0 ? I : 0) * sizeof(float);
  return bytes + 1024;
}

struct SimtWs {
  int64_t* am_off;
  int64_t* lm_off;
  float* hidden;
  float* logits;
  float* dhid;
  int64_t chunk;
};

static SimtWs carve(void* ws, int64_t M, int V, int I) {
  SimtWs w;
  joiner_simt_workspace_bytes(M, V, I, &w.chunk);
  char* p = (char*)ws;
  w.am_off = (int64_t*)p; p += (size_t)M * sizeof(int64_t);
  w.lm_off = (int64_t*)p; p += (size_t)M * sizeof(int64_t);
  w.hidden = (float*)p; p += (size_t)M * (I > 0 ? I : 0) * sizeof(float);
  w.logits = (float*)p; p += (size_t)w.chunk * V * sizeof(float);
  w.dhid = (float*)p;
  return w;
}

static int compute_chunk_logits(const JoinerProblem& p, const SimtWs& w, int64_t row0, int64_t rows,
                                bool compute_hidden, cudaStream_t stream) {
  if (p.I > 0) {
    if (compute_hidden) {
      JointRowOperand a{p.am, p.lm, w.am_off, w.lm_off, row0, p.act};
      StridedOperand b{p.W1, 0, p.V, 1};
      BiasStoreEpilogue ep{w.hidden, p.b1, row0, p.I};
      if (int rc = launch_sgemm<true, true>(1, (int)rows, p.I, p.V, 1, a, b, ep, stream, "joiner_hidden_gemm")) return rc;
    }
    StridedOperand a{w.hidden + row0 * p.I, 0, p.I, 1};
    StridedOperand b{p.W2, 0, p.I, 1};
    BiasStoreEpilogue ep{w.logits, p.b2, 0, p.V};
    return launch_sgemm<true, true>(1, (int)rows, p.V, p.I, 1, a, b, ep, stream, "joiner_logits_gemm");
  }
  int64_t n = rows * p.V;
  ProfScope prof("joint_act_kernel", stream);
  joint_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p.am, p.lm, w.am_off, w.lm_off, row0, rows,
                                                                     p.V, p.act, w.logits);
  return check_launch("joint_act_kernel");
}

int joiner_simt_forward(const JoinerProblem& p, void* workspace, float* lse, float* px, float* py,
                        cudaStream_t stream) {
  const int64_t M = (int64_t)p.B * p.T * p.R;
  if (M == 0) return 0;
  SimtWs w = carve(workspace, M, p.V, p.I);
  {
    ProfScope prof("row_offsets_kernel", stream);
    row_offsets_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(p.ranges, M, p.T, p.R, p.S, p.V, w.am_off,
                                                                         w.lm_off);
  }
  if (int rc = check_launch("row_offsets_kernel")) return rc;
  for (int64_t row0 = 0; row0 < M; row0 += w.chunk) {
    int64_t rows = (M - row0 < w.chunk) ? (M - row0) : w.chunk;
    if (int rc = compute_chunk_logits(p, w, row0, rows, true, stream)) return rc;
    if (int rc = lse_gather_rows(w.logits, p.sym, p.ranges, p.boundary, row0, rows, p.T, p.R, p.V, p.S, p.blank,
                                 p.delay_penalty, lse, px, py, stream))
      return rc;
  }
  return 0;
}

int joiner_simt_materialize(const JoinerProblem& p, void* workspace, float* logits_out, cudaStream_t stream) {
  const int64_t M = (int64_t)p.B * p.T * p.R;
  if (M == 0) return 0;
  SimtWs w = carve(workspace, M, p.V, p.I);
  row_offsets_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(p.ranges, M, p.T, p.R, p.S, p.V, w.am_off,
                                                                       w.lm_off);
  if (int rc = check_launch("row_offsets_kernel")) return rc;
  for (int64_t row0 = 0; row0 < M; row0 += w.chunk) {
    int64_t rows = (M - row0 < w.chunk) ? (M - row0) : w.chunk;
    if (int rc = compute_chunk_logits(p, w, row0, rows, true, stream)) return rc;
    cudaMemcpyAsync(logits_out + row0 * p.V, w.logits, (size_t)rows * p.V * sizeof(float),
                    cudaMemcpyDeviceToDevice, stream);
  }
  return check_launch("joiner_simt_materialize");
}

// Gradients are ACCUMULATED into d_am, d_lm, dW1, db1, dW2, db2 (caller zero-fills or pre-loads).
int joiner_simt_backward(const JoinerProblem& p, void* workspace, const float* lse, const float* occ_px,
                         const float* occ_py, const float* coef, float clamp, float* d_am, float* d_lm,
                         float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream) {
  const int64_t M = (int64_t)p.B * p.T * p.R;
  if (M == 0) return 0;
  SimtWs w = carve(workspace, M, p.V, p.I);
  for (int64_t row0 = 0; row0 < M; row0 += w.chunk) {
    int64_t rows = (M - row0 < w.chunk) ? (M - row0) : w.chunk;
    // hidden is still in the workspace from the forward call
    if (int rc = compute_chunk_logits(p, w, row0, rows, false, stream)) return rc;
    if (int rc = logits_grad_rows(w.logits, p.sym, p.ranges, lse, occ_px, occ_py, coef, row0, rows, p.T, p.R,
                                  p.V, p.S, p.blank, clamp, stream))
      return rc;
    const float* G = w.logits;  // (rows, V) d loss / d logits
    if (p.I > 0) {
      const int splits = (int)((rows + 2047) / 2048) < 64 ? (int)((rows + 2047) / 2048) : 64;
      {  // dhid[m, i] = sum_v G[m, v] W2[v, i]
        StridedOperand a{G, 0, p.V, 1};
        StridedOperand b{p.W2, 0, 1, p.I};
        BiasStoreEpilogue ep{w.dhid, nullptr, 0, p.I};
        if (int rc = launch_sgemm<true, false>(1, (int)rows, p.I, p.V, 1, a, b, ep, stream, "joiner_dhidden_gemm")) return rc;
      }
      {  // dW2[v, i] += sum_m G[m, v] hidden[m, i]
        StridedOperand a{G, 0, 1, p.V};
        StridedOperand b{w.hidden + row0 * p.I, 0, 1, p.I};
        AtomicAccEpilogue ep{dW2, p.I};
        if (int rc = launch_sgemm<false, false>(1, p.V, p.I, (int)rows, splits, a, b, ep, stream, "joiner_dW2_gemm")) return rc;
      }
      {
        ProfScope prof("col_sum_kernel", stream, 2);
        dim3 grid((p.V + 127) / 128, (unsigned)((rows + 255) / 256));
        col_sum_kernel<<<grid, 128, 0, stream>>>(G, rows, p.V, p.V, db2);
        dim3 grid2((p.I + 127) / 128, (unsigned)((rows + 255) / 256));
        col_sum_kernel<<<grid2, 128, 0, stream>>>(w.dhid, rows, p.I, p.I, db1);
        if (int rc = check_launch("col_sum_kernel")) return rc;
      }
      {  // dW1[i, c] += sum_m dhid[m, i] J[m, c]
        StridedOperand a{w.dhid, 0, 1, p.I};
        JointColOperand b{p.am, p.lm, w.am_off, w.lm_off, row0, p.act};
        AtomicAccEpilogue ep{dW1, p.V};
        if (int rc = launch_sgemm<false, false>(1, p.I, p.V, (int)rows, splits, a, b, ep, stream, "joiner_dW1_gemm")) return rc;
      }
      {  // dJ[m, c] = (sum_i dhid[m, i] W1[i, c]) * act'(.)  -> d_am, d_lm
        StridedOperand a{w.dhid, 0, p.I, 1};
        StridedOperand b{p.W1, 0, 1, p.V};
        JointGradEpilogue ep{p.am, p.lm, w.am_off, w.lm_off, row0, p.act, d_am, d_lm};
        if (int rc = launch_sgemm<true, false>(1, (int)rows, p.V, p.I, 1, a, b, ep, stream, "joiner_djoint_gemm")) return rc;
      }
    } else {
      int64_t n = rows * p.V;
      ProfScope prof("joint_grad_kernel", stream);
      joint_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(G, p.am, p.lm, w.am_off, w.lm_off, row0,
                                                                          rows, p.V, p.act, d_am, d_lm);
      if (int rc = check_launch("joint_grad_kernel")) return rc;
    }
  }
  return 0;
}

}  // namespace s2t
