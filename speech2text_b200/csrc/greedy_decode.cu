// Batched greedy RNN-T decoding, device resident.
//
// Replaces the per-utterance Python loop of RnntGreedyDecoding.decode
// (/root/reference/model/decoding.py:225-271, driven by batch_search :32-48 from the validation step's WER metric,
// model/utils.py:115-136): per frame it calls joiner.streaming_step (project, add, activation, out-projection,
// log-softmax: model/joiner/joiner.py:184-207), takes the argmax, and on a non-blank calls predictor.streaming_step
// (model/predictor/stateless_predictor.py:107-124) -- ~10 kernel launches and one .item() host synchronisation per
// lattice step, one utterance at a time.
//
// Here one CTA walks one utterance's lattice from the first frame to the last without leaving the GPU; all
// utterances of the batch run side by side.  The encoder-side projection am = enc W_e^T + b_e is one batched GEMM
// before the kernel (every frame is visited at least once).  Per step the CTA evaluates
//     logits = W2 (W1 act(am[t] + lm) + b1) + b2        (or act(am[t] + lm) without the out-projection)
// as warp-per-row matrix-vector products over L2-resident fp32 weights, and the argmax (log-softmax is monotone, the
// reference's argmax over log-probabilities picks the same class).  On an emission it refreshes
//     lm = W_p (W_o conv(emb[last C tokens]) + b_o) + b_p
// Bound: weight bytes per step through L2 -> SM (I (2V) 4 bytes with the out-projection), i.e. ~10 us per step at
// V = 500, against ~1 ms per step of launches and synchronisation in the reference.
#include "../../include/s2t_b200.h"
#include "common.cuh"

namespace s2t {
namespace {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kDecMaxContext = 8;

// out[j] = bias[j] + sum_k W[j, k] x[k], j < rows: one warp per row, lanes along k (512 contiguous bytes per load)
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float* __restrict__ bias, const float* x, int rows,
                                       int cols, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  for (int j = warp; j < rows; j += kDecWarps) {
    const float* w = W + (int64_t)j * cols;
    float acc = 0.f;
    if (vec) {
      for (int k = lane * 4; k < cols; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        acc = fmaf(wv.x, xv.x, acc);
        acc = fmaf(wv.y, xv.y, acc);
        acc = fmaf(wv.z, xv.z, acc);
        acc = fmaf(wv.w, xv.w, acc);
      }
    } else {
      for (int k = lane; k < cols; k += 32) acc = fmaf(__ldg(w + k), x[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[j] = acc + (bias ? __ldg(bias + j) : 0.f);
  }
}

struct DecodeArgs {
  const float* am;          // (B, T, V)
  const int64_t* lengths;   // (B)
  const float *emb, *conv_w, *Wo, *bo;  // predictor: (N, E), (E, C), (D, E), (D)
  const float *Wp, *bp;                 // joiner _pre_proj: (V, D), (V)
  const float *W1, *b1, *W2, *b2;       // out-projection (I, V), (I), (V, I), (V) or null
  int B, T, V, N, E, C, D, I, act, blank, max_token_step, max_out;
  int64_t* tokens;  // (B, max_out)
  int* n_tokens;    // (B)
};

__global__ void __launch_bounds__(kDecThreads) rnnt_greedy_decode_kernel(const DecodeArgs a) {
  extern __shared__ __align__(16) float sm[];
  // shared vectors (each padded to a multiple of 4 floats): lm[V] joint[V] logits[V] hid[I] h[E] pred[D] + reduction scratch
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  float* lm = sm;
  float* joint = lm + pad4(a.V);
  float* logits = joint + pad4(a.V);
  float* hid = logits + pad4(a.V);
  float* h = hid + pad4(a.I > 0 ? a.I : 1);
  float* pred = h + pad4(a.E);
  float* red_v = pred + pad4(a.D);
  int* red_i = reinterpret_cast<int*>(red_v + kDecWarps);
  __shared__ int ctx[kDecMaxContext];
  __shared__ int s_best;

  const int b = blockIdx.x;
  const int Tb = (int)min((int64_t)a.T, max((int64_t)0, a.lengths[b]));
  const float* am_b = a.am + (int64_t)b * a.T * a.V;
  int64_t* out = a.tokens + (int64_t)b * a.max_out;
  if (threadIdx.x < a.C) ctx[threadIdx.x] = a.blank;  // init state [blank .. blank] + the first input token <blank>
  __syncthreads();

  auto refresh_lm = [&]() {
    // h = depthwise conv over the embeddings of the last C tokens; pred = Wo h + bo; lm = Wp pred + bp
    for (int e = threadIdx.x; e < a.E; e += kDecThreads) {
      float acc = 0.f;
      for (int k = 0; k < a.C; ++k) {
        const int tok = min(max(ctx[k], 0), a.N - 1);
        acc = fmaf(__ldg(a.conv_w + (int64_t)e * a.C + k), __ldg(a.emb + (int64_t)tok * a.E + e), acc);
      }
      h[e] = acc;
    }
    __syncthreads();
    matvec(a.Wo, a.bo, h, a.D, a.E, pred);
    __syncthreads();
    matvec(a.Wp, a.bp, pred, a.V, a.D, lm);
    __syncthreads();
  };

  refresh_lm();
  int t = 0, n_step = 0, n_out = 0;
  while (t < Tb) {
    const float* am_t = am_b + (int64_t)t * a.V;
    for (int v = threadIdx.x; v < a.V; v += kDecThreads) joint[v] = act_fwd(__ldg(am_t + v) + lm[v], a.act);
    __syncthreads();
    const float* scores = joint;
    if (a.I > 0) {
      matvec(a.W1, a.b1, joint, a.I, a.V, hid);
      __syncthreads();
      matvec(a.W2, a.b2, hid, a.V, a.I, logits);
      __syncthreads();
      scores = logits;
    }
    // argmax, first maximum on ties (torch.argmax)
    float bv = kNegInf;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < a.V; v += kDecThreads) {
      const float x = scores[v];
      if (x > bv) {
        bv = x;
        bi = v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      red_v[threadIdx.x >> 5] = bv;
      red_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = red_v[0];
      int fi = red_i[0];
      for (int w = 1; w < kDecWarps; ++w)
        if (red_v[w] > fv || (red_v[w] == fv && red_i[w] < fi)) {
          fv = red_v[w];
          fi = red_i[w];
        }
      s_best = fi;
    }
    __syncthreads();
    const int best = s_best;
    if (best == a.blank || n_step > a.max_token_step) {
      // blank, or the per-frame emission limit: next frame (decoding.py:252-258)
      ++t;
      n_step = 0;
    } else {
      // emit: the lattice moves up, the predictor sees the new token (decoding.py:259-267)
      ++n_step;
      if (threadIdx.x == 0) {
        if (n_out < a.max_out) out[n_out] = best;
        for (int k = 0; k + 1 < a.C; ++k) ctx[k] = ctx[k + 1];
        ctx[a.C - 1] = best;
      }
      ++n_out;
      __syncthreads();
      refresh_lm();
    }
  }
  if (threadIdx.x == 0) a.n_tokens[b] = n_out < a.max_out ? n_out : a.max_out;
}


// ------------------------------------------------------------------------------------------------
// Beam search (RnntBeamDecoding, /root/reference/model/decoding.py:295-425): at most one token per frame.  Per frame
// every beam proposes its cutoff_top_k best classes of log_softmax(joiner(enc_t, pred_beam)); a blank keeps the
// hypothesis, any other class extends it; the beam_size best candidates by accumulated log-probability survive, in the
// order Python's stable sort gives them (descending score, ties in order of creation); hypotheses are NOT merged.
// One CTA per utterance.  The token histories are kept as back-pointers (parent beam, token) per frame and read back
// from the best final beam; the lm rows of the live beams sit in a double-buffered global scratch (L2 resident).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxBeam = 8, kMaxTopK = 8;

struct BeamArgs {
  DecodeArgs d;
  int beam, top_k;
  float* lm_ws;        // (B, 2, beam, V)
  uint8_t* bp_parent;  // (B, T, beam)
  int* bp_tok;         // (B, T, beam)
  float* best_score;   // (B)
};

// block-wide arg-max of v[0..n) (first maximum on ties); every thread gets the result.  red_v / red_i: kDecWarps entries
__device__ __forceinline__ void block_argmax(const float* v, int n, float* red_v, int* red_i, float& out_v, int& out_i) {
  float bv = kNegInf;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += kDecThreads) {
    const float x = v[i];
    if (x > bv) {
      bv = x;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    red_v[threadIdx.x >> 5] = bv;
    red_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  float fv = red_v[0];
  int fi = red_i[0];
  for (int w = 1; w < kDecWarps; ++w)
    if (red_v[w] > fv || (red_v[w] == fv && red_i[w] < fi)) {
      fv = red_v[w];
      fi = red_i[w];
    }
  __syncthreads();
  out_v = fv;
  out_i = fi;
}

__device__ __forceinline__ float block_sum(float x, float* red_v) {
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = x;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < kDecWarps; ++w) s += red_v[w];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(kDecThreads) rnnt_beam_decode_kernel(const BeamArgs g) {
  const DecodeArgs& a = g.d;
  extern __shared__ __align__(16) float sm[];
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  float* joint = sm;
  float* logits = joint + pad4(a.V);
  float* hid = logits + pad4(a.V);
  float* h = hid + pad4(a.I > 0 ? a.I : 1);
  float* pred = h + pad4(a.E);
  float* red_v = pred + pad4(a.D);
  int* red_i = reinterpret_cast<int*>(red_v + kDecWarps);
  __shared__ int ctx[2][kMaxBeam][kDecMaxContext];
  __shared__ float score[2][kMaxBeam];
  __shared__ float c_score[kMaxBeam * kMaxTopK];
  __shared__ int c_parent[kMaxBeam * kMaxTopK], c_tok[kMaxBeam * kMaxTopK];
  __shared__ int s_parent[kMaxBeam], s_tok[kMaxBeam], s_n;

  const int b = blockIdx.x;
  const int Tb = (int)min((int64_t)a.T, max((int64_t)0, a.lengths[b]));
  const float* am_b = a.am + (int64_t)b * a.T * a.V;
  float* lm_b = g.lm_ws + (int64_t)b * 2 * g.beam * a.V;
  uint8_t* bpp = g.bp_parent + (int64_t)b * a.T * g.beam;
  int* bpt = g.bp_tok + (int64_t)b * a.T * g.beam;
  const int top_k = min(g.top_k, a.V);

  // lm row of a context: h = depthwise conv over the embeddings of the last C tokens; pred = Wo h + bo; lm = Wp pred + bp
  auto lm_of = [&](const int* c, float* out) {
    for (int e = threadIdx.x; e < a.E; e += kDecThreads) {
      float acc = 0.f;
      for (int k = 0; k < a.C; ++k) {
        const int tok = min(max(c[k], 0), a.N - 1);
        acc = fmaf(__ldg(a.conv_w + (int64_t)e * a.C + k), __ldg(a.emb + (int64_t)tok * a.E + e), acc);
      }
      h[e] = acc;
    }
    __syncthreads();
    matvec(a.Wo, a.bo, h, a.D, a.E, pred);
    __syncthreads();
    matvec(a.Wp, a.bp, pred, a.V, a.D, out);
    __syncthreads();
  };

  int cur = 0, n_beams = 1;
  if (threadIdx.x < a.C) ctx[0][0][threadIdx.x] = a.blank;
  if (threadIdx.x == 0) score[0][0] = 0.f;
  __syncthreads();
  lm_of(ctx[0][0], lm_b);
  __threadfence_block();

  for (int t = 0; t < Tb; ++t) {
    const float* am_t = am_b + (int64_t)t * a.V;
    // ---- candidates of every beam ----
    for (int i = 0; i < n_beams; ++i) {
      const float* lm_i = lm_b + ((int64_t)cur * g.beam + i) * a.V;
      for (int v = threadIdx.x; v < a.V; v += kDecThreads) joint[v] = act_fwd(__ldg(am_t + v) + lm_i[v], a.act);
      __syncthreads();
      float* sc = joint;
      if (a.I > 0) {
        matvec(a.W1, a.b1, joint, a.I, a.V, hid);
        __syncthreads();
        matvec(a.W2, a.b2, hid, a.V, a.I, logits);
        __syncthreads();
        sc = logits;
      }
      float mx;
      int mi;
      block_argmax(sc, a.V, red_v, red_i, mx, mi);
      float part = 0.f;
      for (int v = threadIdx.x; v < a.V; v += kDecThreads) part += __expf(sc[v] - mx);
      const float lse = mx + __logf(block_sum(part, red_v));
      for (int k = 0; k < top_k; ++k) {
        float kv;
        int ki;
        if (k == 0) {
          kv = mx;
          ki = mi;
        } else {
          block_argmax(sc, a.V, red_v, red_i, kv, ki);
        }
        if (threadIdx.x == 0) {
          c_score[i * top_k + k] = score[cur][i] + (kv - lse);
          c_parent[i * top_k + k] = i;
          c_tok[i * top_k + k] = ki;
          sc[ki] = kNegInf;  // out of the next passes
        }
        __syncthreads();
      }
    }
    // ---- the beam_size best candidates: descending score, ties in order of creation (Python's stable sort) ----
    if (threadIdx.x == 0) {
      const int nc = n_beams * top_k;
      const int keep = min(g.beam, nc);
      bool used[kMaxBeam * kMaxTopK];
      for (int c = 0; c < nc; ++c) used[c] = false;
      for (int r = 0; r < keep; ++r) {
        int best = -1;
        for (int c = 0; c < nc; ++c)
          if (!used[c] && (best < 0 || c_score[c] > c_score[best])) best = c;
        used[best] = true;
        s_parent[r] = c_parent[best];
        s_tok[r] = c_tok[best];
        score[cur ^ 1][r] = c_score[best];
        bpp[(int64_t)t * g.beam + r] = (uint8_t)c_parent[best];
        bpt[(int64_t)t * g.beam + r] = c_tok[best];
      }
      s_n = keep;
    }
    __syncthreads();
    const int n_new = s_n;
    // ---- the survivors' predictor state: unchanged after a blank, one step further otherwise ----
    for (int r = 0; r < n_new; ++r) {
      const int p = s_parent[r], tok = s_tok[r];
      float* dst = lm_b + ((int64_t)(cur ^ 1) * g.beam + r) * a.V;
      if (threadIdx.x < a.C) {
        int c;
        if (tok == a.blank) c = ctx[cur][p][threadIdx.x];
        else c = threadIdx.x + 1 < a.C ? ctx[cur][p][threadIdx.x + 1] : tok;
        ctx[cur ^ 1][r][threadIdx.x] = c;
      }
      __syncthreads();
      if (tok == a.blank) {
        const float* src = lm_b + ((int64_t)cur * g.beam + p) * a.V;
        for (int v = threadIdx.x; v < a.V; v += kDecThreads) dst[v] = src[v];
        __syncthreads();
      } else {
        lm_of(ctx[cur ^ 1][r], dst);
      }
    }
    __threadfence_block();
    __syncthreads();
    cur ^= 1;
    n_beams = n_new;
  }
  // ---- read the best hypothesis back through the back-pointers ----
  if (threadIdx.x == 0) {
    int n = 0;
    int i = 0;
    for (int t = Tb - 1; t >= 0; --t) {
      if (bpt[(int64_t)t * g.beam + i] != a.blank) ++n;
      i = bpp[(int64_t)t * g.beam + i];
    }
    a.n_tokens[b] = n;
    g.best_score[b] = Tb > 0 ? score[cur][0] : 0.f;
    int64_t* out = a.tokens + (int64_t)b * a.max_out;
    i = 0;
    int w = n;
    for (int t = Tb - 1; t >= 0; --t) {
      const int tok = bpt[(int64_t)t * g.beam + i];
      if (tok != a.blank && --w < a.max_out) out[w] = tok;
      i = bpp[(int64_t)t * g.beam + i];
    }
  }
}

}  // namespace
}  // namespace s2t

using namespace s2t;

extern "C" {

int s2t_rnnt_greedy_decode(const float* am, const int64_t* lengths, const float* emb, const float* conv_w, const float* Wo,
                           const float* bo, const float* Wp, const float* bp, const float* W1, const float* b1,
                           const float* W2, const float* b2, int B, int T, int V, int N, int E, int C, int D, int I, int act,
                           int blank, int max_token_step, int max_out, int64_t* tokens, int* n_tokens, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B > 0 && T > 0 && V > 0 && N > 0 && E > 0 && D > 0 && C >= 1 && C <= kDecMaxContext,
              "rnnt_greedy_decode: bad dims B=%d T=%d V=%d N=%d E=%d C=%d D=%d", B, T, V, N, E, C, D);
  S2T_REQUIRE(I == 0 || (W1 && b1 && W2 && b2), "rnnt_greedy_decode: out-projection weights missing");
  S2T_REQUIRE(blank >= 0 && blank < V && max_out > 0, "rnnt_greedy_decode: blank %d / max_out %d", blank, max_out);
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  const size_t smem = (size_t)(3 * pad4(V) + pad4(I > 0 ? I : 1) + pad4(E) + pad4(D) + 2 * kDecWarps + 8) * sizeof(float);
  S2T_REQUIRE(smem <= 200 * 1024, "rnnt_greedy_decode: V=%d too large for the per-utterance shared-memory vectors", V);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rnnt_greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DecodeArgs a{am, lengths, emb, conv_w, Wo, bo, Wp, bp, W1, b1, W2, b2, B, T, V, N, E, C, D, I, act, blank,
               max_token_step, max_out, tokens, n_tokens};
  ProfScope prof("rnnt_greedy_decode_kernel", st);
  rnnt_greedy_decode_kernel<<<B, kDecThreads, smem, st>>>(a);
  return check_launch("rnnt_greedy_decode_kernel");
}

size_t s2t_rnnt_beam_workspace_bytes(int B, int T, int V, int beam) {
  return (size_t)B * 2 * beam * V * sizeof(float) + (size_t)B * T * beam * (sizeof(int) + 1) + 256;
}

int s2t_rnnt_beam_decode(const float* am, const int64_t* lengths, const float* emb, const float* conv_w, const float* Wo,
                         const float* bo, const float* Wp, const float* bp, const float* W1, const float* b1,
                         const float* W2, const float* b2, int B, int T, int V, int N, int E, int C, int D, int I, int act,
                         int blank, int beam, int top_k, void* workspace, int64_t* tokens, int* n_tokens, float* best_score,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(B > 0 && T > 0 && V > 0 && N > 0 && E > 0 && D > 0 && C >= 1 && C <= kDecMaxContext,
              "rnnt_beam_decode: bad dims B=%d T=%d V=%d N=%d E=%d C=%d D=%d", B, T, V, N, E, C, D);
  S2T_REQUIRE(I == 0 || (W1 && b1 && W2 && b2), "rnnt_beam_decode: out-projection weights missing");
  S2T_REQUIRE(blank >= 0 && blank < V, "rnnt_beam_decode: blank %d", blank);
  S2T_REQUIRE(beam >= 1 && beam <= kMaxBeam && top_k >= 1 && top_k <= kMaxTopK, "rnnt_beam_decode: beam %d (<= %d), top_k %d (<= %d)",
              beam, kMaxBeam, top_k, kMaxTopK);
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  const size_t smem = (size_t)(2 * pad4(V) + pad4(I > 0 ? I : 1) + pad4(E) + pad4(D) + 2 * kDecWarps + 8) * sizeof(float);
  S2T_REQUIRE(smem <= 200 * 1024, "rnnt_beam_decode: V=%d too large for the per-utterance shared-memory vectors", V);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rnnt_beam_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BeamArgs g;
  g.d = DecodeArgs{am, lengths, emb, conv_w, Wo, bo, Wp, bp, W1, b1, W2, b2, B, T, V, N, E, C, D, I, act, blank, 1, T, tokens, n_tokens};
  g.beam = beam;
  g.top_k = top_k;
  char* w = (char*)workspace;
  g.lm_ws = (float*)w;
  w += (size_t)B * 2 * beam * V * sizeof(float);
  g.bp_tok = (int*)w;
  w += (size_t)B * T * beam * sizeof(int);
  g.bp_parent = (uint8_t*)w;
  g.best_score = best_score;
  ProfScope prof("rnnt_beam_decode_kernel", st);
  rnnt_beam_decode_kernel<<<B, kDecThreads, smem, st>>>(g);
  return check_launch("rnnt_beam_decode_kernel");
}

}  // extern "C"
