// CTC loss with the log-softmax fused in.
//
// Replaces, for the CTC branch of PrunedRnntTask / CtcHybridRnnt
// (/root/reference/task_factory/rnnt_task.py:341-349, 485-496),
//   F.log_softmax(logits, -1).transpose(0, 1).float() + nn.CTCLoss(blank, reduction, zero_infinity)
//   /root/reference/model/loss/ctc_loss.py:35-41
// without ever writing the (T, B, V) log-probabilities: three kernels, all HBM / latency bound.
//
//   ctc_emit_kernel     one warp per frame (b, t < T_b): lse over V (online max / sum, 16-byte loads) and the
//                       S_b + 1 emissions the lattice can use, E[b, t, 0] = logit[blank] - lse,
//                       E[b, t, j] = logit[label_{j-1}] - lse        (reads the logits once: V*4 bytes per frame)
//   ctc_lattice_kernel  alpha and beta over the 2 S_b + 1 states, one CTA per (utterance, direction) -- the two
//                       directions run side by side on different SMs.  One state per thread, the previous step's
//                       values handed over through a double-buffered shared-memory row (one barrier per step),
//                       the next frame's emission prefetched under the barrier.  Values are kept relative to an
//                       fp64 offset that is re-based (block-wide max) every 8 frames: exp(alpha + beta - log P) at
//                       |log P| in the thousands otherwise carries ~1e-3 relative error in fp32.
//   ctc_grad_kernel     one warp per frame: grad[b, t, v] = coef_b (softmax(logits)[v] - sum_{s: l'_s = v} gamma_t(s)),
//                       gamma_t(s) = exp(alpha_t(s) + beta~_t(s) - log P) (beta~ excludes the emission at t); zero for
//                       padding frames and (zero_infinity) for utterances without a valid alignment.
//                       (reads the logits once more and writes the gradient once: 2 V*4 bytes per frame)
//
// Algorithmic bytes per utterance: 3 T V 4 (two logits reads, one gradient write) + 5 T (S+1) 4 (E written and read
// twice, alpha / beta~ written and read: (2S+1) ~ 2 (S+1) states each).
#include <stdlib.h>

#include "../../include/s2t_b200.h"
#include "common.cuh"

namespace s2t {
namespace {

constexpr int kRebase = 8;  // frames between re-basings of the running offset

struct CtcWs {
  float* E;        // (B, T, S+1)
  float* alpha;    // (B, T, L)  L = 2 S + 1, relative to off_a[b, t]
  float* beta;     // (B, T, L)  beta~ (emission at t excluded), relative to off_b[b, t]
  double* off_a;   // (B, T)
  double* off_b;   // (B, T)
  int* first;      // (B, S): first label position that carries the same class as position i
  size_t bytes;
};

CtcWs ctc_carve(void* ws, int B, int T, int S) {
  CtcWs w;
  char* p = (char*)ws;
  auto take = [&](size_t n) {
    char* q = p;
    p += (n + 255) / 256 * 256;
    return q;
  };
  const size_t L = 2 * (size_t)S + 1;
  w.E = (float*)take((size_t)B * T * (S + 1) * sizeof(float));
  w.alpha = (float*)take((size_t)B * T * L * sizeof(float));
  w.beta = (float*)take((size_t)B * T * L * sizeof(float));
  w.off_a = (double*)take((size_t)B * T * sizeof(double));
  w.off_b = (double*)take((size_t)B * T * sizeof(double));
  w.first = (int*)take((size_t)B * (S > 0 ? S : 1) * sizeof(int));
  w.bytes = (size_t)(p - (char*)ws);
  return w;
}

__device__ __forceinline__ int clamp_len(int64_t v, int hi) { return (int)(v < 0 ? 0 : (v > hi ? hi : v)); }

// ---- emissions ----------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(256) ctc_emit_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                                       const int64_t* __restrict__ in_len,
                                                       const int64_t* __restrict__ tgt_len, int B, int T, int S, int V,
                                                       int blank, float* __restrict__ lse_out, float* __restrict__ E) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / T), t = (int)(row % T);
  const int Tb = clamp_len(in_len[b], T);
  if (t >= Tb) {
    if (lane == 0) lse_out[row] = 0.f;
    return;
  }
  const float* x = logits + row * V;
  float mx = kNegInf, sum = 0.f;
  if (kVec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int n4 = V >> 2;
    // four 16-byte loads in flight per lane
    for (int i = lane; i < n4; i += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + 32 * u < n4) ? __ldg(x4 + i + 32 * u) : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
      float m4 = mx;
#pragma unroll
      for (int u = 0; u < 4; ++u) m4 = fmaxf(m4, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
      if (m4 > mx) {
        sum *= __expf(mx - m4);  // (-inf - finite) -> 0 on the first pass
        mx = m4;
      }
      if (mx > kNegInf) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sum += __expf(v[u].x - mx) + __expf(v[u].y - mx) + __expf(v[u].z - mx) + __expf(v[u].w - mx);
      }
    }
  } else {
    for (int i = lane; i < V; i += 32) {
      const float v = __ldg(x + i);
      if (v > mx) {
        sum *= __expf(mx - v);
        mx = v;
      }
      if (mx > kNegInf) sum += __expf(v - mx);
    }
  }
  // combine the 32 (max, sum) pairs
  const float wm = warp_max(mx);
  sum = (mx > kNegInf) ? sum * __expf(mx - wm) : 0.f;
  sum = warp_sum(sum);
  const float lse = wm + logf(sum);
  if (lane == 0) lse_out[row] = lse;
  const int Sb = clamp_len(tgt_len[b], S);
  float* e = E + row * (S + 1);
  for (int j = lane; j <= Sb; j += 32) {
    const int c = j == 0 ? blank : (int)targets[(int64_t)b * S + j - 1];
    e[j] = (c >= 0 && c < V) ? __ldg(x + c) - lse : kNegInf;
  }
}

// ---- lattice ------------------------------------------------------------------------------------
__device__ __forceinline__ float log_add3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// block-wide max over the states (values of threads without a state are -inf); red = 33 floats of shared memory
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float m = (lane < nw) ? red[lane] : kNegInf;
  m = warp_max(m);
  __syncthreads();
  return m;
}

// (A warp-per-direction variant -- eight states per lane in registers, neighbours by shuffle, no barrier -- measured
// 0.33 ms against 0.25 for this kernel at BASELINE config 4, 0.44 with the emissions of eight frames in registers: a lone
// warp does not overlap its eight log-adds well enough to beat one state per thread plus a barrier.)
// blockIdx.x = 2 b + direction (0: alpha, forward in time; 1: beta~, backward).  Threads own the states
// s = tid, tid + blockDim, ... (one each for L <= blockDim).  Shared memory: 2 rows of (L + 2) floats + 33.
__global__ void __launch_bounds__(1024) ctc_lattice_kernel(const float* __restrict__ E, const int64_t* __restrict__ targets,
                                                           const int64_t* __restrict__ in_len,
                                                           const int64_t* __restrict__ tgt_len, int B, int T, int S,
                                                           float* __restrict__ alpha, float* __restrict__ beta,
                                                           double* __restrict__ off_a, double* __restrict__ off_b,
                                                           float* __restrict__ nll, int* __restrict__ first) {
  extern __shared__ float sm[];
  const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
  const int Tb = clamp_len(in_len[b], T), Sb = clamp_len(tgt_len[b], S);
  const int L = 2 * Sb + 1, Lfull = 2 * S + 1;
  float* row[2] = {sm + 2, sm + 2 + (Lfull + 2) + 2};  // two guard cells in front of each row (s - 1, s - 2 reads)
  float* red = sm + 2 * (Lfull + 4);
  if (Tb == 0) {
    if (dir == 0 && threadIdx.x == 0) nll[b] = Sb == 0 ? 0.f : INFINITY;
    return;
  }
  const int64_t* lab = targets + (int64_t)b * S;
  if (dir == 0) {
    // labels repeat: the gradient kernel sums the occupations of a class in the slot of its first position
    for (int i = threadIdx.x; i < Sb; i += blockDim.x) {
      int f = i;
      const int64_t c = lab[i];
      for (int q = 0; q < i; ++q)
        if (lab[q] == c) {
          f = q;
          break;
        }
      first[(int64_t)b * S + i] = f;
    }
  }
  const float* Eb = E + (int64_t)b * T * (S + 1);
  float* out = (dir == 0 ? alpha : beta) + (int64_t)b * T * Lfull;
  double* off = (dir == 0 ? off_a : off_b) + (int64_t)b * T;
  for (int i = threadIdx.x; i < 2 * (Lfull + 4); i += blockDim.x) sm[i] = kNegInf;
  __syncthreads();
  double base = 0.0;  // running offset (replicated in every thread)
  if (L <= (int)blockDim.x) {
    // One state per thread (every target the reference's configs produce: 2 S + 1 <= 1024).  tau counts the steps
    // of this direction, t = tau (alpha) or T_b - 1 - tau (beta~); rows are indexed so that the two neighbours a
    // state reads are the cells in front of it (alpha: r = s; beta: r = L - 1 - s).  The emissions of the next four
    // steps are already in registers: the recursion itself never waits for global memory.
    const int s = threadIdx.x;
    const bool active = s < L;
    const int j = (s & 1) ? (s + 1) >> 1 : 0;
    const int r = dir == 0 ? s : L - 1 - s;
    bool skip = false;
    if (active && (s & 1)) {
      if (dir == 0) skip = s >= 3 && lab[(s - 1) >> 1] != lab[(s - 3) >> 1];
      else skip = s + 2 < L && lab[(s - 1) >> 1] != lab[(s + 1) >> 1];
    }
    const float* ep = Eb + j;
    auto frame = [&](int tau) { return dir == 0 ? tau : Tb - 1 - tau; };
    float eq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) eq[u] = (active && u < Tb) ? __ldg(ep + (int64_t)frame(u) * (S + 1)) : 0.f;
    for (int tau0 = 0; tau0 < Tb; tau0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tau = tau0 + u;
        if (tau >= Tb) break;
        const int t = frame(tau);
        const float e = eq[u];
        if (active && tau + 4 < Tb) eq[u] = __ldg(ep + (int64_t)frame(tau + 4) * (S + 1));
        float* cur = row[tau & 1];
        const float* prev = row[(tau & 1) ^ 1];
        float v = kNegInf;
        if (active) {
          float stored;
          if (tau == 0) {
            if (dir == 0) {
              v = s <= 1 ? e : kNegInf;
              stored = v;
            } else {
              stored = s >= L - 2 ? 0.f : kNegInf;
              v = stored + e;
            }
          } else {
            const float a3 = log_add3(prev[r], prev[r - 1], skip ? prev[r - 2] : kNegInf);
            v = a3 + e;
            stored = dir == 0 ? v : a3;
          }
          cur[r] = v;
          out[(int64_t)t * Lfull + s] = stored;
        }
        if (threadIdx.x == 0) off[t] = base;  // the frame's stored values are relative to the base before a re-basing
        if ((tau % kRebase) == kRebase - 1 && tau + 1 < Tb) {
          const float m = block_max(v, red);  // includes the barrier that publishes cur
          if (m > kNegInf) {
            if (active) cur[r] -= m;
            base += (double)m;
          }
        }
        __syncthreads();
      }
    }
    if (dir == 0 && threadIdx.x == 0) {
      const float* last = row[(Tb - 1) & 1];
      const float a = last[L - 1], c = L >= 2 ? last[L - 2] : kNegInf;
      const float m = fmaxf(a, c);
      const double lp = m == kNegInf ? -INFINITY : (double)(m + __logf(__expf(a - m) + __expf(c - m))) + base;
      nll[b] = (float)(-lp);
    }
    return;
  }
  // general path: several states per thread
  if (dir == 0) {
    // alpha_t(s) = e_t(s) + logadd(alpha_{t-1}(s), alpha_{t-1}(s-1), [skip] alpha_{t-1}(s-2))
    for (int t = 0; t < Tb; ++t) {
      float* cur = row[t & 1];
      const float* prev = row[(t & 1) ^ 1];
      float vmax = kNegInf;
      for (int s = threadIdx.x; s < L; s += blockDim.x) {
        const int j = (s & 1) ? (s + 1) >> 1 : 0;
        const float e = Eb[(int64_t)t * (S + 1) + j];
        float v;
        if (t == 0) {
          v = s <= 1 ? e : kNegInf;
        } else {
          const bool skip = (s & 1) && s >= 3 && lab[(s - 1) >> 1] != lab[(s - 3) >> 1];
          v = log_add3(prev[s], prev[s - 1], skip ? prev[s - 2] : kNegInf) + e;
        }
        cur[s] = v;
        out[(int64_t)t * Lfull + s] = v;
        vmax = fmaxf(vmax, v);
      }
      if (threadIdx.x == 0) off[t] = base;
      if ((t % kRebase) == kRebase - 1 && t + 1 < Tb) {
        const float m = block_max(vmax, red);  // includes the barrier that publishes cur
        if (m > kNegInf) {
          for (int s = threadIdx.x; s < L; s += blockDim.x) cur[s] -= m;
          base += (double)m;
        }
        // out[] of this frame was stored relative to the old base (recorded in off[t]); later frames use the new one
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float* last = row[(Tb - 1) & 1];
      // when frame Tb-1 itself was a rebase frame the row is relative to the base recorded BEFORE the rebase
      const float a = last[L - 1], c = L >= 2 ? last[L - 2] : kNegInf;
      const float m = fmaxf(a, c);
      const double lp = m == kNegInf ? -INFINITY : (double)(m + __logf(__expf(a - m) + __expf(c - m))) + base;
      nll[b] = (float)(-lp);
    }
  } else {
    // beta_t(s) = e_t(s) + beta~_t(s),  beta~_t(s) = logadd(beta_{t+1}(s), beta_{t+1}(s+1), [skip] beta_{t+1}(s+2));
    // rows hold beta (with the emission); the stored quantity is beta~.  Guard cells sit BEHIND the row here:
    // rows are indexed from the end (index L - 1 - s), so that s + 1 / s + 2 are the two cells in front.
    for (int t = Tb - 1; t >= 0; --t) {
      float* cur = row[t & 1];
      const float* prev = row[(t & 1) ^ 1];
      float vmax = kNegInf;
      for (int s = threadIdx.x; s < L; s += blockDim.x) {
        const int j = (s & 1) ? (s + 1) >> 1 : 0;
        const float e = Eb[(int64_t)t * (S + 1) + j];
        const int r = L - 1 - s;  // reversed index: r - 1 <-> s + 1, r - 2 <-> s + 2
        float bt;
        if (t == Tb - 1) {
          bt = s >= L - 2 ? 0.f : kNegInf;
        } else {
          const bool skip = (s & 1) && s + 2 < L && lab[(s - 1) >> 1] != lab[(s + 1) >> 1];
          bt = log_add3(prev[r], prev[r - 1], skip ? prev[r - 2] : kNegInf);
        }
        out[(int64_t)t * Lfull + s] = bt;
        const float v = bt + e;
        cur[r] = v;
        vmax = fmaxf(vmax, v);
      }
      if (threadIdx.x == 0) off[t] = base;
      if (((Tb - 1 - t) % kRebase) == kRebase - 1 && t > 0) {
        const float m = block_max(vmax, red);
        if (m > kNegInf) {
          for (int s = threadIdx.x; s < L; s += blockDim.x) cur[L - 1 - s] -= m;
          base += (double)m;
        }
      }
      __syncthreads();
    }
  }
}

// ---- gradient -----------------------------------------------------------------------------------
// One warp per frame.  Dense pass: coef * softmax -> grad row (16-byte stores).  Sparse pass: the occupations of the
// frame's states are summed per distinct class in shared memory (labels repeat) and subtracted from the row; the
// row belongs to this warp alone, so plain read-modify-writes after a __syncwarp() are enough.
template <bool kVec>
__global__ void __launch_bounds__(256) ctc_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                                       const int64_t* __restrict__ in_len,
                                                       const int64_t* __restrict__ tgt_len, const float* __restrict__ lse,
                                                       const float* __restrict__ alpha, const float* __restrict__ beta,
                                                       const double* __restrict__ off_a, const double* __restrict__ off_b,
                                                       const float* __restrict__ nll, const float* __restrict__ grad_nll,
                                                       const int* __restrict__ first, int zero_infinity, int B, int T, int S, int V, int blank,
                                                       float* __restrict__ grad) {
  extern __shared__ float acc_all[];  // per warp: S + 1 accumulators (slot 0 blank, slot j the j-th label's class)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t frame = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (frame >= (int64_t)B * T) return;
  const int b = (int)(frame / T), t = (int)(frame % T);
  const int Tb = clamp_len(in_len[b], T), Sb = clamp_len(tgt_len[b], S);
  float* g = grad + frame * V;
  const float nl = nll[b];
  const bool dead = t >= Tb || (zero_infinity && !isfinite(nl));
  const float cf = dead ? 0.f : grad_nll[b];
  if (dead || cf == 0.f) {
    if (kVec) {
      float4* g4 = reinterpret_cast<float4*>(g);
      for (int i = lane; i < (V >> 2); i += 32) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = lane; i < V; i += 32) g[i] = 0.f;
    }
    return;
  }
  const float* x = logits + frame * V;
  const float l = lse[frame];
  if (kVec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* g4 = reinterpret_cast<float4*>(g);
    const int n4 = V >> 2;
    for (int i = lane; i < n4; i += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + 32 * u < n4) ? __ldg(x4 + i + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + 32 * u < n4)
          g4[i + 32 * u] = make_float4(cf * __expf(v[u].x - l), cf * __expf(v[u].y - l), cf * __expf(v[u].z - l),
                                       cf * __expf(v[u].w - l));
    }
  } else {
    for (int i = lane; i < V; i += 32) g[i] = cf * __expf(__ldg(x + i) - l);
  }
  // occupations.  An infinite nll without zero_infinity gives NaN gradients in torch as well (inf - inf).
  const int L = 2 * Sb + 1, Lfull = 2 * S + 1;
  const float k = (float)(off_a[frame] + off_b[frame] + (double)nl);  // alpha + beta~ - log P, log P = -nll
  const float* a = alpha + frame * Lfull;
  const float* bt = beta + frame * Lfull;
  const int64_t* lab = targets + (int64_t)b * S;
  float* acc = acc_all + warp * (S + 1);
  for (int j = lane; j <= Sb; j += 32) acc[j] = 0.f;
  __syncwarp();
  float blank_sum = 0.f;
  for (int s = lane; s < L; s += 32) {
    const float gm = __expf(a[s] + bt[s] + k);
    if (s & 1) {
      // the slot of the first position that carries the same class collects the class's occupation
      atomicAdd(&acc[first[(int64_t)b * S + ((s - 1) >> 1)] + 1], gm);
    } else {
      blank_sum += gm;
    }
  }
  blank_sum = warp_sum(blank_sum);
  __syncwarp();
  // the dense pass of this warp is complete and visible to all of its lanes after the barrier above
  for (int j = lane; j <= Sb; j += 32) {
    const float v = j == 0 ? blank_sum : acc[j];
    if (v == 0.f) continue;
    const int c = j == 0 ? blank : (int)lab[j - 1];
    if (j > 0 && c == blank) {  // a label equal to the blank class (never produced by the reference's tokenizer)
      atomicAdd(g + c, -cf * v);
      continue;
    }
    if (c >= 0 && c < V) {
      if (j == 0) {
        // labels equal to blank (handled above with atomics) may touch this address too
        atomicAdd(g + c, -cf * v);
      } else {
        g[c] -= cf * v;
      }
    }
  }
}

}  // namespace
}  // namespace s2t

using namespace s2t;

extern "C" {

size_t s2t_ctc_workspace_bytes(int B, int T, int S, int V) {
  (void)V;
  return ctc_carve(nullptr, B, T, S).bytes + 256;
}

int s2t_ctc_loss_fwd(const float* logits, const int64_t* targets, const int64_t* logit_lengths,
                     const int64_t* target_lengths, int B, int T, int S, int V, int blank, void* workspace, float* lse,
                     float* nll, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B > 0 && T > 0 && S >= 0 && V > 0, "ctc_loss: bad dims B=%d T=%d S=%d V=%d", B, T, S, V);
  S2T_REQUIRE(blank >= 0 && blank < V, "ctc_loss: blank %d out of range", blank);
  CtcWs w = ctc_carve(workspace, B, T, S);
  const int64_t frames = (int64_t)B * T;
  {
    ProfScope prof("ctc_emit_kernel", st);
    const unsigned grid = (unsigned)((frames + 7) / 8);
    const bool vec = (V % 4 == 0) && ((uintptr_t)logits % 16 == 0);
    if (vec) ctc_emit_kernel<true><<<grid, 256, 0, st>>>(logits, targets, logit_lengths, target_lengths, B, T, S, V, blank, lse, w.E);
    else ctc_emit_kernel<false><<<grid, 256, 0, st>>>(logits, targets, logit_lengths, target_lengths, B, T, S, V, blank, lse, w.E);
  }
  if (int rc = check_launch("ctc_emit_kernel")) return rc;
  {
    ProfScope prof("ctc_lattice_kernel", st);
    const int Lfull = 2 * S + 1;
    int threads = ((Lfull + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 64) threads = 64;
    const size_t smem = (size_t)(2 * (Lfull + 4) + 40) * sizeof(float);
    S2T_REQUIRE(smem <= 200 * 1024, "ctc_loss: target length %d too long for the lattice kernel's shared memory", S);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(ctc_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ctc_lattice_kernel<<<2 * B, threads, smem, st>>>(w.E, targets, logit_lengths, target_lengths, B, T, S, w.alpha, w.beta,
                                                     w.off_a, w.off_b, nll, w.first);
  }
  return check_launch("ctc_lattice_kernel");
}

int s2t_ctc_loss_bwd(const float* logits, const int64_t* targets, const int64_t* logit_lengths,
                     const int64_t* target_lengths, int B, int T, int S, int V, int blank, const void* workspace,
                     const float* lse, const float* nll, const float* grad_nll, int zero_infinity, float* grad_logits,
                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CtcWs w = ctc_carve(const_cast<void*>(workspace), B, T, S);
  const int64_t frames = (int64_t)B * T;
  ProfScope prof("ctc_grad_kernel", st);
  // warps per block bounded by the shared-memory accumulators (S + 1 floats per warp)
  int warps = 8;
  while (warps > 1 && (size_t)warps * (S + 1) * sizeof(float) > 40 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * (S + 1) * sizeof(float);
  S2T_REQUIRE(smem <= 48 * 1024, "ctc_loss: target length %d too long for the gradient kernel's shared memory", S);
  const unsigned grid = (unsigned)((frames + warps - 1) / warps);
  const bool vec = (V % 4 == 0) && (((uintptr_t)logits | (uintptr_t)grad_logits) % 16 == 0);
  if (vec)
    ctc_grad_kernel<true><<<grid, warps * 32, smem, st>>>(logits, targets, logit_lengths, target_lengths, lse, w.alpha, w.beta,
                                                          w.off_a, w.off_b, nll, grad_nll, w.first, zero_infinity, B, T, S, V, blank,
                                                          grad_logits);
  else
    ctc_grad_kernel<false><<<grid, warps * 32, smem, st>>>(logits, targets, logit_lengths, target_lengths, lse, w.alpha, w.beta,
                                                           w.off_a, w.off_b, nll, grad_nll, w.first, zero_infinity, B, T, S, V, blank,
                                                           grad_logits);
  return check_launch("ctc_grad_kernel");
}

}  // extern "C"
