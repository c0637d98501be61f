// Tensor-core (bf16 operands, fp32 accumulate in TMEM) pruned joiner + loss front end.
//
// Same contract as joiner_simt.cu -- replaces k2.do_rnnt_pruning, the add/activation/out-projection
// of /root/reference/model/joiner/joiner.py:121-123, 176-178 and the logsumexp + gather of
// k2.rnnt_loss_pruned (/root/reference/model/loss/pruned_rnnt_loss.py:39-48), plus their gradients
//  -- but every contraction runs on tcgen05 through tc_gemm.cuh:
//
//   forward   hidden = act(am + lm[ranges]) W1^T + b1      A built on the fly (never in HBM)
//             logits = hidden W2^T + b2  -> per-tile (max, sum-exp) + sym/blank gather in the
//             epilogue; the (B,T,R,V) logits never leave TMEM/registers
//   backward  G = d loss / d logits is recomputed tile by tile from hidden and lse, then
//             dhidden = G W2, dW2 = G^T hidden, dJ = (dhidden W1) * act', dW1 = dhidden^T act(.)
//
// Row-chunking bounds the only (B,T,R,V)-sized scratch (the bf16 G of one chunk, kept in both
// orientations for the two contractions that consume it).
// TODO(perf): chain logits -> G -> dhidden inside one kernel so G stays in smem.
#include <stdlib.h>
#include <limits.h>
#include <mutex>
#include <unordered_map>
#include <deque>
#include "joiner.cuh"
#include "tc_gemm.cuh"

#include <cuda.h>

namespace s2t {
namespace {

using namespace tc;

// CTA pairs (tcgen05 cta_group::2) remain for the producer-fed hidden contraction of the no-keep path; the bulk-fed
// contractions walk the live-tile list (padding frames skipped), which is a single-CTA schedule.
constexpr int kPair = S2T_PAIR;

constexpr float kLog2e = 1.4426950408889634f;

// CTA tile width (columns of the TMEM accumulator) and smem ring depth of the joiner contractions.
// 256 columns = both 256-column TMEM accumulator buffers of the persistent kernel; 48 KB per stage.
// (128-wide tiles were measured slower: the on-the-fly A operand is then produced once per 128
// instead of once per 256 output columns.)
constexpr int kBN = 256;
constexpr int kNStages = 4;   // 192 KB ring
// contractions over the inner dimension (K = Ip <= 256: logits, their gradient, dh) keep B resident in shared memory
constexpr int kResSteps = 4;    // resident k-steps of B: 4 x 32 KB
constexpr int kNStagesRes = 4;  // A-only ring of 16 KB stages next to it
constexpr int kLseGroups = S2T_BULK_EPI_GROUPS;  // epilogue groups of the (bulk-fed) logits -> LSE kernel

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float act_fwd_fast(float x, int act) { return act == kRelu ? fmaxf(x, 0.f) : tanh_fast(x); }
__device__ __forceinline__ float act_bwd_fast(float x, int act) {
  if (act == kRelu) return x > 0.f ? 1.f : 0.f;
  float y = tanh_fast(x);
  return 1.f - y * y;
}

// lane l ends up with sum over the warp of v[l]   (31 shuffles)
__device__ __forceinline__ float warp_column_sums(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = hi ? v[i] : v[i + off];
      float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---- per-row metadata -----------------------------------------------------------------------
// am_row[m] = b * T + t and lm_row[m] = b * (S+1) + s(t, r): the rows of am / lm that joiner row m adds up
__global__ void tc_row_meta_kernel(const int64_t* __restrict__ ranges, const int64_t* __restrict__ sym,
                                   int64_t rows, int T, int R, int S, int blank, int* __restrict__ am_row,
                                   int* __restrict__ lm_row, int* __restrict__ row_sym) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  int64_t bt = m / R;
  int r = (int)(m % R);
  int64_t b = bt / T;
  int s = ranges ? (int)ranges[m] : r;
  int sc = min(max(s, 0), S);
  am_row[m] = (int)bt;
  lm_row[m] = (int)(b * (S + 1) + sc);
  if (row_sym) row_sym[m] = (sym && s >= 0 && s < S) ? (int)sym[b * S + s] : blank;
}

// ---- A producers ----------------------------------------------------------------------------
// Both producers build act(am[m, v] + lm[m, v]) -> bf16 straight into the swizzled stage.  Sixteen lanes
// cover the 64 vocabulary entries (256 contiguous bytes of am and of lm) that one joiner row contributes
// to a k-step, so every warp-wide load reads whole 128-byte lines; a k-step is pipelined in two halves
// of four rows per thread, the loads of the next half in flight while the current one is converted.
struct JointQuad {
  float4 a[4], l[4];
};

// loads entries v .. v+3 of the four rows (am_row[j], lm_row[j]); a negative row index marks a dead row
__device__ __forceinline__ void joint_load_quad(JointQuad& q, const float* am, const float* lm, const int (&ar)[4],
                                                const int (&lr)[4], int v, int V, bool vec) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool live = ar[j] >= 0;
    const float* a = am + (int64_t)(live ? ar[j] : 0) * V + v;
    const float* l = lm + (int64_t)(live ? lr[j] : 0) * V + v;
    if (live && vec && v + 4 <= V) {
      q.a[j] = __ldg(reinterpret_cast<const float4*>(a));
      q.l[j] = __ldg(reinterpret_cast<const float4*>(l));
    } else {
      float xa[4], xl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = live && (v + e < V);
        xa[e] = ok ? __ldg(a + e) : 0.f;
        xl[e] = ok ? __ldg(l + e) : 0.f;
      }
      q.a[j] = make_float4(xa[0], xa[1], xa[2], xa[3]);
      q.l[j] = make_float4(xl[0], xl[1], xl[2], xl[3]);
    }
  }
}

// 4 entries -> 8 bytes.  Padding (v >= V or dead row) must come out as exact zeros: act(0) = 0 for relu and tanh.
__device__ __forceinline__ uint2 joint_act4(const float4& a, const float4& l, int act) {
  return make_uint2(pack_bf16x2(act_fwd_fast(a.x + l.x, act), act_fwd_fast(a.y + l.y, act)),
                    pack_bf16x2(act_fwd_fast(a.z + l.z, act), act_fwd_fast(a.w + l.w, act)));
}

// K-major A: block rows = joiner rows m of the tile, K = vocabulary.  Thread (warp w, lane l) serves rows
// 2w + l/16 + 16 i (i = 0..7) and the four entries 4 (l % 16) .. +3 of every 64-entry k-step.
struct JointRowProducer {
  static constexpr bool kBulk = false;
  const float* am;
  const float* lm;
  const int* am_row;
  const int* lm_row;
  int64_t M;
  int V, act;
  // optional by-product: the stage image is also the packed operand block (row tile, k-step) of
  // J = act(am + lm[ranges]); kept for the weight-gradient contraction of the backward pass
  uint8_t* Jp;
  int j_row_blocks;
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int c = lane & 15, rbase = warp * 2 + (lane >> 4);
    const bool vec = ((V & 3) == 0);
    int ar[2][4], lr[2][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = (int64_t)pc.m_tile * 128 + rbase + 16 * i;
      ar[i >> 2][i & 3] = m < M ? __ldg(am_row + m) : -1;
      lr[i >> 2][i & 3] = m < M ? __ldg(lm_row + m) : 0;
    }
    // byte offset inside the stage of this thread's 8 bytes in row rbase (+ 16 i rows: the swizzle only sees row & 7)
    const int off = rbase * 128 + ((((c >> 1) ^ (rbase & 7)) & 7) << 4) + (c & 1) * 8;
    JointQuad q0, q1;
    joint_load_quad(q0, am, lm, ar[0], lr[0], pc.ks0 * 64 + c * 4, V, vec);
    for (int it = 0; it < pc.n_it; ++it) {
      const int v = (pc.ks0 + it) * 64 + c * 4;
      joint_load_quad(q1, am, lm, ar[1], lr[1], v, V, vec);
      pc.wait_empty(it);
      uint8_t* dst = pc.stage(it) + off;
      uint8_t* jdst = (Jp != nullptr && pc.n_tile == 0 && pc.valid)
                          ? Jp + packed_block_index(pc.m_tile, pc.ks0 + it, j_row_blocks) * kBlockBytes + off
                          : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 o = joint_act4(q0.a[j], q0.l[j], act);
        *reinterpret_cast<uint2*>(dst + j * 16 * 128) = o;
        if (jdst) *reinterpret_cast<uint2*>(jdst + j * 16 * 128) = o;
      }
      if (it + 1 < pc.n_it) joint_load_quad(q0, am, lm, ar[0], lr[0], v + 64, V, vec);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 o = joint_act4(q1.a[j], q1.l[j], act);
        *reinterpret_cast<uint2*>(dst + (4 + j) * 16 * 128) = o;
        if (jdst) *reinterpret_cast<uint2*>(jdst + (4 + j) * 16 * 128) = o;
      }
      pc.arrive_full(it);
    }
  }
};

// MN-major A: stage = 2 groups x [64 contraction rows (joiner rows m) x 64 vocabulary entries].  A warp
// covers the 128 entries of one contraction row (512 contiguous bytes of am and of lm): lane l owns entries
// 4 l .. +3 (group l / 16), thread (warp w) serves contraction rows w + 8 i (i = 0..7) of every k-step.
struct JointMnProducer {
  static constexpr bool kBulk = false;
  const float* am;
  const float* lm;
  const int* am_row;
  const int* lm_row;
  int64_t row0, M;
  int V, act;
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int g = lane >> 4, c = lane & 15;
    const int v = pc.m_tile * 128 + lane * 4;
    const bool vec = ((V & 3) == 0);
    auto load_rows = [&](int (&ar)[2][4], int (&lr)[2][4], int ks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = row0 + (int64_t)ks * 64 + warp + 8 * i;
        ar[i >> 2][i & 3] = m < M ? __ldg(am_row + m) : -1;
        lr[i >> 2][i & 3] = m < M ? __ldg(lm_row + m) : 0;
      }
    };
    const int off = g * kGroupBytes + warp * 128 + ((((c >> 1) ^ (warp & 7)) & 7) << 4) + (c & 1) * 8;
    int ar[2][4], lr[2][4], ar_n[2][4], lr_n[2][4];
    load_rows(ar, lr, pc.ks0);
    JointQuad q0, q1;
    joint_load_quad(q0, am, lm, ar[0], lr[0], v, V, vec);
    for (int it = 0; it < pc.n_it; ++it) {
      const bool more = it + 1 < pc.n_it;
      if (more) load_rows(ar_n, lr_n, pc.ks0 + it + 1);
      joint_load_quad(q1, am, lm, ar[1], lr[1], v, V, vec);
      pc.wait_empty(it);
      uint8_t* dst = pc.stage(it) + off;
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint2*>(dst + j * 8 * 128) = joint_act4(q0.a[j], q0.l[j], act);
      if (more) joint_load_quad(q0, am, lm, ar_n[0], lr_n[0], v, V, vec);
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint2*>(dst + (4 + j) * 8 * 128) = joint_act4(q1.a[j], q1.l[j], act);
      pc.arrive_full(it);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ar[i >> 2][i & 3] = ar_n[i >> 2][i & 3];
        lr[i >> 2][i & 3] = lr_n[i >> 2][i & 3];
      }
    }
  }
};

// J = act(am + lm[ranges]) for every joiner row -> bf16 packed operand (rows m, cols v).  A CTA owns 32 rows:
// their am / lm row indices are staged once, then every warp converts (row, 128-column segment) tasks, four at
// a time so that their loads are in flight together.  A lane owns four consecutive vocabulary entries, so each
// warp-wide load reads 512 contiguous bytes of one am (lm) row and each 8-byte store instruction fills two whole
// 128-byte block rows.  Rows beyond M and columns beyond V come out as zeros.  Used when J is kept for the
// backward pass: the hidden contraction is then fed by bulk copies like every other one.
constexpr int kJpRows = 32;

// tile_live[i] = 1 when any of the rows 128 i .. 128 i + 127 belongs to a frame t < T_b of its utterance.  Padding
// frames occupy the tail of every utterance's (T R)-row block: at c3 a sixth of the row tiles hold nothing else,
// and no kernel of the path computes, stores or reads them.  live_idx lists the live tiles in ascending order,
// live_prefix[i] counts the live tiles before tile i (Mt + 1 entries).  One block: flags, block scan, lists.
__global__ void __launch_bounds__(1024) tc_live_tiles_kernel(const int64_t* __restrict__ boundary, int64_t M, int T, int R,
                                                            int Mt, uint8_t* __restrict__ tile_live,
                                                            int* __restrict__ live_idx, int* __restrict__ live_prefix) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < Mt; base += 1024) {
    const int i = base + threadIdx.x;
    int live = 0;
    if (i < Mt) {
      const int64_t m0 = (int64_t)i * 128, m1 = min(M, m0 + 128) - 1;
      const int64_t f0 = m0 / R, f1 = m1 / R;  // first and last frame (b T + t) of the tile
      for (int64_t b = f0 / T; b <= f1 / T && !live; ++b) {
        const int ts = (int)(max(f0, b * T) - b * T);  // first frame of the tile inside utterance b
        const int Tb = boundary ? min(max((int)boundary[4 * b + 3], 0), T) : T;
        live = ts < Tb;
      }
      tile_live[i] = (uint8_t)live;
    }
    // exclusive scan of the flags over the block
    const unsigned bal = __ballot_sync(0xffffffffu, live != 0);
    const int in_warp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    const int pos = before + in_warp;
    if (i < Mt) {
      live_prefix[i] = pos;
      if (live) live_idx[pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = pos + (live ? 1 : 0);
    __syncthreads();
  }
  if (threadIdx.x == 0) live_prefix[Mt] = carry;
}

template <int kAct>
__global__ void __launch_bounds__(256) joint_pack_kernel(const float* __restrict__ am, const float* __restrict__ lm,
                                                         const int* __restrict__ am_row, const int* __restrict__ lm_row,
                                                         int64_t M, int V, int row_blocks, int k_blocks,
                                                         uint8_t* __restrict__ Jp, const uint8_t* __restrict__ tile_live) {
  __shared__ int ar[kJpRows], lr[kJpRows];
  const int64_t m0 = (int64_t)blockIdx.x * kJpRows;
  if (tile_live != nullptr && tile_live[m0 >> 7] == 0) return;  // padding frames only: nobody reads this block
  if (threadIdx.x < kJpRows) {
    const int64_t m = m0 + threadIdx.x;
    ar[threadIdx.x] = m < M ? __ldg(am_row + m) : -1;
    lr[threadIdx.x] = m < M ? __ldg(lm_row + m) : 0;
  }
  __syncthreads();
  const bool vec = (V & 3) == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = (int)(m0 >> 7), r_in = (int)(m0 & 127);  // the 32 rows sit in one 128-row block
  const int segs = k_blocks / 2;                           // 128-column segments per row (Vp is a multiple of 256)
  const int tasks = kJpRows * segs;                        // task = seg * 32 + row: a warp walks down the rows
  for (int base = warp; base < tasks; base += 8 * 4) {
    float4 a[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = base + 8 * u;
      const int r = task & (kJpRows - 1), v = (task / kJpRows) * 128 + lane * 4;
      const bool live = task < tasks && ar[r] >= 0 && v < V;
      const float* pa = am + (int64_t)max(ar[r], 0) * V + v;
      const float* pl = lm + (int64_t)lr[r] * V + v;
      if (live && vec && v + 4 <= V) {
        a[u] = __ldg(reinterpret_cast<const float4*>(pa));
        l[u] = __ldg(reinterpret_cast<const float4*>(pl));
      } else {
        float xa[4], xl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = live && v + j < V;
          xa[j] = ok ? __ldg(pa + j) : 0.f;
          xl[j] = ok ? __ldg(pl + j) : 0.f;
        }
        a[u] = make_float4(xa[0], xa[1], xa[2], xa[3]);
        l[u] = make_float4(xl[0], xl[1], xl[2], xl[3]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = base + 8 * u;
      if (task >= tasks) break;
      const int r = task & (kJpRows - 1), seg = task / kJpRows;
      const float x0 = a[u].x + l[u].x, x1 = a[u].y + l[u].y, x2 = a[u].z + l[u].z, x3 = a[u].w + l[u].w;
      uint2 o;
      if (kAct == kRelu) {
        o = make_uint2(pack_bf16x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)), pack_bf16x2(fmaxf(x2, 0.f), fmaxf(x3, 0.f)));
      } else {
        o = make_uint2(pack_bf16x2(tanh_fast(x0), tanh_fast(x1)), pack_bf16x2(tanh_fast(x2), tanh_fast(x3)));
      }
      // lane -> k-block seg * 2 + lane / 16, 16-byte chunk (lane % 16) / 2, half lane % 2
      uint8_t* blk = Jp + packed_block_index(rb, seg * 2 + (lane >> 4), row_blocks) * kBlockBytes;
      *reinterpret_cast<uint2*>(blk + block_chunk_offset(r_in + r, (lane & 15) >> 1) + (lane & 1) * 8) = o;
    }
  }
}

// ---- epilogue helpers -----------------------------------------------------------------------
// write 32 consecutive K-elements (columns n..n+31 of the accumulator) of every row of the warp into a packed
// operand; transposed through shared memory so that four lanes write the 64 contiguous bytes a row owns
// inside its (swizzled) 128-byte block row.  ctx.m is the thread's own row (rows of a warp are consecutive).
__device__ __forceinline__ void store_packed_row32(const EpiCtx& ctx, uint8_t* packed, int row_blocks, int n,
                                                   const float (&x)[32]) {
  uint4 mine[4];
  pack_row32_bf16(x, mine);
  const int64_t m0 = (int64_t)ctx.m - (ctx.t & 31);
  const int kb = n >> 6, c0 = (n & 63) >> 3;
  warp_transposed_chunk_b16(ctx, mine, [&](int r, int q, uint4 v) {
    const int64_t rg = m0 + r;
    uint8_t* blk = packed + packed_block_index((int)(rg >> 7), kb, row_blocks) * kBlockBytes;
    *reinterpret_cast<uint4*>(blk + block_chunk_offset((int)(rg & 127), c0 + q)) = v;
  });
}
// hidden = acc + b1 -> Hp (rows m, cols i)
struct HiddenEpi {
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  const float* b1;
  int I;
  int64_t M;
  uint8_t* Hp;
  int h_row_blocks;
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    float x[32];
    const bool live = ctx.m < M;
    if (n + 32 <= I) {  // interior chunk: one select per row instead of a range test per element
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = live ? acc[j] + __ldg(b1 + n + j) : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = (live && n + j < I) ? acc[j] + __ldg(b1 + n + j) : 0.f;
    }
    store_packed_row32(ctx, Hp, h_row_blocks, n, x);
  }
};

// logits = acc + b2: per-tile (max, sum exp) and the gathered sym / blank logits
struct LseEpi {
  static constexpr int kScratchBytes = 0;
  const float* b2;
  const int* row_sym;
  int V, blank, n_tiles;
  int64_t M;
  float* part;  // (M, n_tiles, 2)
  float* sym_logit;
  float* blank_logit;
  struct State { float mx, sum; int csym; };
  __device__ void begin(State& st, const EpiCtx& ctx) const {
    st.mx = kNegInf;
    st.sum = 0.f;
    st.csym = ctx.m < M ? row_sym[ctx.m] : -1;
  }
  __device__ void chunk(State& st, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    if (ctx.m >= M || n >= V) return;
    float x[32];
    float cm = kNegInf;
    if (n + 32 <= V) {  // interior chunk: no per-element range test (the epilogue is issue-bound)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = acc[j] + __ldg(b2 + n + j);
        cm = fmaxf(cm, x[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = (n + j < V) ? acc[j] + __ldg(b2 + n + j) : kNegInf;
        cm = fmaxf(cm, x[j]);
      }
    }
    if (cm > st.mx) {
      st.sum *= ex2_approx((st.mx - cm) * kLog2e);  // (-inf - finite) -> 0 on the first chunk
      st.mx = cm;
    }
    const float nm = -st.mx * kLog2e;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {  // arguments <= 0; padding columns hold -inf -> 0
      s0 += ex2_approx(fmaf(x[j], kLog2e, nm));
      s1 += ex2_approx(fmaf(x[j + 1], kLog2e, nm));
    }
    st.sum += s0 + s1;
    if (st.csym >= n && st.csym < n + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j == st.csym) sym_logit[ctx.m] = x[j];
    }
    if (blank >= n && blank < n + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j == blank) blank_logit[ctx.m] = x[j];
    }
  }
  __device__ void end(State& st, const EpiCtx& ctx) const {
    if (ctx.m >= M) return;
    float* p = part + ((int64_t)ctx.m * n_tiles + ctx.part) * 2;  // n_tiles counts (n_tile, group) partials
    p[0] = st.mx;
    p[1] = st.sum;
  }
};

__global__ void lse_combine_kernel(const float* __restrict__ part, const float* __restrict__ sym_logit,
                                   const float* __restrict__ blank_logit, const int64_t* __restrict__ boundary,
                                   int64_t rows, int n_tiles, int T, int R, float delay_penalty,
                                   float* __restrict__ lse, float* __restrict__ px, float* __restrict__ py) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  if (boundary) {  // padding frame: its row tile may never have been computed; the lattice ignores these entries
    const int64_t bt = m / R;
    const int b = (int)(bt / T), t = (int)(bt % T);
    if (t >= min(max((int)boundary[4 * b + 3], 0), T)) {
      lse[m] = 0.f;
      px[m] = 0.f;
      py[m] = 0.f;
      return;
    }
  }
  const float* p = part + m * n_tiles * 2;
  float mx = kNegInf;
  for (int i = 0; i < n_tiles; ++i) mx = fmaxf(mx, p[2 * i]);
  float s = 0.f;
  for (int i = 0; i < n_tiles; ++i) s += p[2 * i + 1] * expf(p[2 * i] - mx);
  const float l = mx + logf(s);
  lse[m] = l;
  float xv = sym_logit[m] - l;
  if (delay_penalty != 0.f) {
    const int b = (int)(m / ((int64_t)T * R));
    const int t = (int)((m / R) % T);
    const int Tb = boundary ? (int)boundary[4 * b + 3] : T;
    xv += delay_penalty * (0.5f * (float)(Tb - 1) - (float)t);
  }
  px[m] = xv;
  py[m] = blank_logit[m] - l;
}

// G = coef * clip(occ_px [v == sym] + occ_py [v == blank] - (occ_px + occ_py) softmax) for a row chunk
struct GradEpi {
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  const float* b2;
  const int* row_sym;
  const float* lse;
  const float* occ_px;
  const float* occ_py;
  const float* coef;  // per utterance
  int64_t row0, M;
  int TR, V, blank;
  float clamp;
  uint8_t* Gp;   // rows = chunk-local m, cols = v (Vp / 64 column blocks)
  int g_row_blocks;
  float* db2;
  struct State { float l, ox, oy, cf; int csym; bool live; };
  __device__ void begin(State& st, const EpiCtx& ctx) const {
    const int64_t m = row0 + ctx.m;
    st.live = m < M;
    st.l = 0.f; st.ox = 0.f; st.oy = 0.f; st.cf = 0.f; st.csym = -1;
    if (st.live) {
      st.l = lse[m];
      st.ox = occ_px[m];
      st.oy = occ_py[m];
      st.cf = coef[m / TR];
      st.csym = row_sym[m];
      if (st.ox + st.oy == 0.f || st.cf == 0.f) st.live = false;
    }
    if (!st.live) {
      st.ox = 0.f;
      st.oy = 0.f;
    }
  }
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State& st, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    float x[32];
    // softmax term: -cf (occ_px + occ_py) exp(logit - lse); dead rows use -inf so that it is exactly 0
    const float gneg = st.live ? -(st.ox + st.oy) * st.cf : 0.f;
    const float nl2 = st.live ? -st.l * kLog2e : kNegInf;
    const float oxc = st.ox * st.cf, oyc = st.oy * st.cf;
    if (clamp > 0.f) {
      // torchaudio's gradient clamp acts on the per-utterance gradient before the upstream scale
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int v = n + j;
        float val = 0.f;
        if (st.live && v < V) {
          val = -(st.ox + st.oy) * exp2f(fmaf(acc[j] + __ldg(b2 + v), kLog2e, nl2));
          if (v == st.csym) val += st.ox;
          if (v == blank) val += st.oy;
          val = fminf(fmaxf(val, -clamp), clamp) * st.cf;
        }
        x[j] = val;
      }
    } else if (n + 32 <= V) {
      // whole chunk inside the vocabulary (all but the last one): no per-element range test, the bias folded into
      // the exponent offset, the blank column fixed up under a warp-uniform branch -- the epilogue is issue-bound
      const int cj = st.csym - n;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = fmaf(__ldg(b2 + n + j), kLog2e, nl2);
        float val = gneg * ex2_approx(fmaf(acc[j], kLog2e, t));
        if (j == cj) val += oxc;
        x[j] = val;
      }
      if (blank >= n && blank < n + 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j == blank) x[j] += oyc;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int v = n + j;
        const float b = (v < V) ? __ldg(b2 + v) : 0.f;
        float val = gneg * ex2_approx(fmaf(acc[j] + b, kLog2e, nl2));
        if (v == st.csym) val += oxc;
        if (v == blank) val += oyc;
        x[j] = (v < V) ? val : 0.f;
      }
    }
    store_packed_row32(ctx, Gp, g_row_blocks, n, x);
    const float cs = warp_column_sums(x);
    const int lane = threadIdx.x & 31;
    if (n + lane < V && cs != 0.f) atomicAdd(db2 + n + lane, cs);
  }
  // no kColSums: measured 0.0705 -> 0.074 ms with the shared-memory sums here (this epilogue is issue-bound and its
  // global reductions are fire-and-forget), whereas the dhidden epilogue below gained 0.058 -> 0.0455
};

// dhidden -> DHp (rows chunk-local m, cols i), db1
struct DHiddenEpi {
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  int I;
  uint8_t* DHp;
  int dh_row_blocks;
  float* db1;
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    float x[32];
    if (n + 32 <= I) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = acc[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = (n + j < I) ? acc[j] : 0.f;
    }
    store_packed_row32(ctx, DHp, dh_row_blocks, n, x);
    const float cs = warp_column_sums(x);
    const int lane = threadIdx.x & 31;
    if (n + lane < I && cs != 0.f) {
      if (ctx.colsum) atomicAdd(ctx.colsum + (n - ctx.tile_col0) + lane, cs);
      else atomicAdd(db1 + n + lane, cs);
    }
  }
  static constexpr bool kColSums = true;
  __device__ void flush_colsum(int col, float v) const {
    if (col < I) atomicAdd(db1 + col, v);
  }
};

// dh = dhidden W1 (before the activation derivative), bf16 row-major (chunk rows, Vp): the GEMM epilogue
// is a plain store; the segmented reductions into d_am / d_lm run as two fully parallel kernels below.
struct StoreRowsBf16Epi {
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  __nv_bfloat16* out;
  int ld;
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    uint4 mine[4];
    pack_row32_bf16(acc, mine);
    const int64_t m0 = (int64_t)ctx.m - (ctx.t & 31);
    warp_transposed_chunk_b16(ctx, mine, [&](int r, int q, uint4 v) {
      *reinterpret_cast<uint4*>(out + (m0 + r) * ld + n + q * 8) = v;
    });
  }
};

// dJ = dh * act'(am + lm[ranges]) reduced both ways in one pass over dh:
//   d_am[b,t,v] += sum_r dJ[(b,t,r), v]                       (A.5: sum over the band slots of a frame)
//   d_lm[b,s,v] += sum over the (t, r) with sb[t] + r = s of dJ[(b,t,r), v]
// One CTA owns kFrames consecutive frames of one utterance; every thread owns four vocabulary columns, so
// the frame sums live in registers and the per-symbol-position sums in thread-private columns of a
// shared-memory tile (no barrier, no shared atomics).  The band of kFrames frames spans only a few symbol
// positions; the tile is flushed to d_lm with 16-byte reductions, since neighbouring frame chunks meet in
// the same rows.  Wider spans (unpruned lattices) are processed in windows of kMaxSpan positions.
constexpr int kDjFrames = 8;  // 16 left 2.2 waves of CTAs at c3 (three rounds of 28 us): wave quantisation, not bandwidth, set the time
constexpr int kDjMaxSpan = 16;
constexpr int kDjCols = 512;  // columns per pass: 128 threads x 4

template <bool kVec, int kRB>
__global__ void __launch_bounds__(128) djoint_reduce_kernel(const __nv_bfloat16* __restrict__ dh, int ld,
                                                            const float* __restrict__ am, const float* __restrict__ lm,
                                                            const int64_t* __restrict__ ranges,
                                                            const int64_t* __restrict__ boundary, int64_t row0,
                                                            int64_t rows, int64_t M, int T, int S, int R, int V, int act,
                                                            bool am_accumulate, float* __restrict__ d_am,
                                                            float* __restrict__ d_lm) {
  extern __shared__ float acc[];  // [kDjMaxSpan][kDjCols]
  __shared__ int sbs[kDjFrames];
  const int b = blockIdx.y, t0 = blockIdx.x * kDjFrames;
  const int Tb = boundary ? min((int)boundary[4 * b + 3], T) : T;  // padding frames carry no gradient
  const int t1 = min(t0 + kDjFrames, Tb);
  if (!am_accumulate) {
    // padding frames carry no gradient: their d_am rows are zero-filled here (the buffer is not memset beforehand:
    // every live frame below is first written with a plain store)
    for (int t = max(t0, Tb); t < min(t0 + kDjFrames, T); ++t) {
      float* drow = d_am + ((int64_t)b * T + t) * V;
      for (int v = threadIdx.x; v < V; v += blockDim.x) drow[v] = 0.f;
    }
  }
  if (t0 >= t1) return;
  const int64_t bt0 = (int64_t)b * T;
  const int64_t row_end = min(row0 + rows, M);
  if ((bt0 + t1) * R <= row0 || (bt0 + t0) * R >= row_end) return;  // no row of this chunk
  if (threadIdx.x < t1 - t0) sbs[threadIdx.x] = ranges ? (int)ranges[(bt0 + t0 + threadIdx.x) * R] : 0;
  __syncthreads();
  const int s_lo = min(max(sbs[0], 0), S), s_hi = min(max(sbs[t1 - t0 - 1] + R - 1, 0), S);
  float* mine = acc + threadIdx.x * 4;
  for (int w_lo = s_lo; w_lo <= s_hi; w_lo += kDjMaxSpan) {
    const int w_hi = min(w_lo + kDjMaxSpan - 1, s_hi);
    for (int vb = 0; vb < V; vb += kDjCols) {
      const int v = vb + threadIdx.x * 4;
      if (v >= V) continue;
      for (int i = 0; i <= w_hi - w_lo; ++i) *reinterpret_cast<float4*>(mine + i * kDjCols) = make_float4(0.f, 0.f, 0.f, 0.f);
      // Common case: the CTA's frames lie inside this row chunk and their band fits one window.  Then every
      // slot of every frame is live, the dh rows of a frame are R consecutive rows and its lm rows R consecutive
      // symbol positions: no predicates, no index arithmetic beyond two running pointers.  The loads of frame
      // t+1 are in flight while frame t is reduced (two register sets used alternately).
      const bool lean = kVec && R == kRB && s_hi - s_lo < kDjMaxSpan && (bt0 + t0) * R >= row0 &&
                        (bt0 + t1) * R <= row_end && sbs[t1 - t0 - 1] + R - 1 <= S && sbs[0] >= 0;
      if (lean) {
        struct Frame {
          uint2 g[kRB];
          float4 l[kRB];
          float4 a;
        };
        const __nv_bfloat16* dhp = dh + ((bt0 + t0) * R - row0) * ld + v;
        const float* amp = am + (bt0 + t0) * V + v;
        const float* lmb = lm + (int64_t)b * (S + 1) * V + v;
        const int64_t frame_stride = (int64_t)R * ld;
        auto load_frame = [&](Frame& f, int i) {  // i = t - t0
          const __nv_bfloat16* gp = dhp + i * frame_stride;
          const float* lp = lmb + (int64_t)sbs[i] * V;
          f.a = __ldg(reinterpret_cast<const float4*>(amp + (int64_t)i * V));
#pragma unroll
          for (int q = 0; q < kRB; ++q) {
            f.g[q] = *reinterpret_cast<const uint2*>(gp + (int64_t)q * ld);
            f.l[q] = __ldg(reinterpret_cast<const float4*>(lp + (int64_t)q * V));
          }
        };
        auto reduce_frame = [&](const Frame& f, int i) {
          float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
          float* cell0 = mine + (sbs[i] - w_lo) * kDjCols;
#pragma unroll
          for (int q = 0; q < kRB; ++q) {
            float4* cell = reinterpret_cast<float4*>(cell0 + q * kDjCols);
            float4 c = *cell;
            const float x0 = __uint_as_float(f.g[q].x << 16) * act_bwd_fast(f.a.x + f.l[q].x, act);
            const float x1 = __uint_as_float(f.g[q].x & 0xffff0000u) * act_bwd_fast(f.a.y + f.l[q].y, act);
            const float x2 = __uint_as_float(f.g[q].y << 16) * act_bwd_fast(f.a.z + f.l[q].z, act);
            const float x3 = __uint_as_float(f.g[q].y & 0xffff0000u) * act_bwd_fast(f.a.w + f.l[q].w, act);
            ds.x += x0; ds.y += x1; ds.z += x2; ds.w += x3;
            c.x += x0; c.y += x1; c.z += x2; c.w += x3;
            *cell = c;
          }
          float* drow = d_am + (bt0 + t0 + i) * V + v;
          if (am_accumulate) {
            const float4 old = *reinterpret_cast<float4*>(drow);
            ds.x += old.x; ds.y += old.y; ds.z += old.z; ds.w += old.w;
          }
          *reinterpret_cast<float4*>(drow) = ds;
        };
        const int nf = t1 - t0;
        // the two register sets keep the loads of one frame in flight, which is less than a DRAM round trip of
        // work: the dh / am lines of the frames after that are pulled into L2 ahead of their loads
        auto prefetch_frame = [&](int i) {
          if (i < nf) {
            const __nv_bfloat16* gp = dhp + i * frame_stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(amp + (int64_t)i * V));
#pragma unroll
            for (int q = 0; q < kRB; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + (int64_t)q * ld));
          }
        };
        prefetch_frame(2);
        prefetch_frame(3);
        Frame fa, fb;
        load_frame(fa, 0);
        for (int i = 0; i < nf; i += 2) {
          if (i + 1 < nf) load_frame(fb, i + 1);
          prefetch_frame(i + 4);
          reduce_frame(fa, i);
          if (i + 1 < nf) {
            if (i + 2 < nf) load_frame(fa, i + 2);
            prefetch_frame(i + 5);
            reduce_frame(fb, i + 1);
          }
        }
      } else if (kVec && ranges == nullptr && (w_hi - w_lo + 1) % kRB == 0 && w_hi <= S && (bt0 + t0) * R >= row0 &&
                 (bt0 + t1) * R <= row_end) {
        // Wide band without pruning (R = S + 1, every frame holds every symbol position): the window's slots are the
        // same kRB-row blocks in every frame, all live -- the frame loop of the lean path over (frame, block) pairs.
        struct Blk {
          uint2 g[kRB];
          float4 l[kRB];
          float4 a;
        };
        const int n_blk = (w_hi - w_lo + 1) / kRB, nf = t1 - t0, n_it = nf * n_blk;
        const __nv_bfloat16* dhp = dh + ((bt0 + t0) * R + w_lo - row0) * ld + v;
        const float* amp = am + (bt0 + t0) * V + v;
        const float* lmb = lm + ((int64_t)b * (S + 1) + w_lo) * V + v;
        const int64_t frame_stride = (int64_t)R * ld;
        auto load_blk = [&](Blk& f, int it) {
          const int i = it / n_blk, q0 = (it - i * n_blk) * kRB;
          const __nv_bfloat16* gp = dhp + i * frame_stride + (int64_t)q0 * ld;
          const float* lp = lmb + (int64_t)q0 * V;
          f.a = __ldg(reinterpret_cast<const float4*>(amp + (int64_t)i * V));
#pragma unroll
          for (int q = 0; q < kRB; ++q) {
            f.g[q] = *reinterpret_cast<const uint2*>(gp + (int64_t)q * ld);
            f.l[q] = __ldg(reinterpret_cast<const float4*>(lp + (int64_t)q * V));
          }
        };
        float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
        auto reduce_blk = [&](const Blk& f, int it) {
          const int i = it / n_blk, blk = it - i * n_blk;
          float* cell0 = mine + blk * kRB * kDjCols;
#pragma unroll
          for (int q = 0; q < kRB; ++q) {
            float4* cell = reinterpret_cast<float4*>(cell0 + q * kDjCols);
            float4 c = *cell;
            const float x0 = __uint_as_float(f.g[q].x << 16) * act_bwd_fast(f.a.x + f.l[q].x, act);
            const float x1 = __uint_as_float(f.g[q].x & 0xffff0000u) * act_bwd_fast(f.a.y + f.l[q].y, act);
            const float x2 = __uint_as_float(f.g[q].y << 16) * act_bwd_fast(f.a.z + f.l[q].z, act);
            const float x3 = __uint_as_float(f.g[q].y & 0xffff0000u) * act_bwd_fast(f.a.w + f.l[q].w, act);
            ds.x += x0; ds.y += x1; ds.z += x2; ds.w += x3;
            c.x += x0; c.y += x1; c.z += x2; c.w += x3;
            *cell = c;
          }
          if (blk == n_blk - 1) {  // the frame's share of this window is complete
            float* drow = d_am + (bt0 + t0 + i) * V + v;
            if (am_accumulate || w_lo > 0) {  // an earlier window already stored this frame's first slots
              const float4 old = *reinterpret_cast<float4*>(drow);
              ds.x += old.x; ds.y += old.y; ds.z += old.z; ds.w += old.w;
            }
            *reinterpret_cast<float4*>(drow) = ds;
            ds = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        auto prefetch_blk = [&](int it) {  // the dh rows of a block a few blocks ahead: into L2 before their loads
          if (it < n_it) {
            const int i = it / n_blk, q0 = (it - i * n_blk) * kRB;
            const __nv_bfloat16* gp = dhp + i * frame_stride + (int64_t)q0 * ld;
#pragma unroll
            for (int q = 0; q < kRB; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + (int64_t)q * ld));
          }
        };
        prefetch_blk(2);
        prefetch_blk(3);
        prefetch_blk(4);
        prefetch_blk(5);
        Blk fa, fb;
        load_blk(fa, 0);
        for (int it = 0; it < n_it; it += 2) {
          if (it + 1 < n_it) load_blk(fb, it + 1);
          prefetch_blk(it + 6);
          reduce_blk(fa, it);
          if (it + 1 < n_it) {
            if (it + 2 < n_it) load_blk(fa, it + 2);
            prefetch_blk(it + 7);
            reduce_blk(fb, it + 1);
          }
        }
      } else {
      for (int t = t0; t < t1; ++t) {
          const int sbt = sbs[t - t0];
          const int r_lo = max(0, w_lo - sbt), r_hi = min(R - 1, w_hi - sbt);
          const int64_t m_base = (bt0 + t) * R;
          if (r_lo > r_hi || m_base + r_hi < row0 || m_base + r_lo >= row_end) continue;
          float a[4], dsum[4] = {0.f, 0.f, 0.f, 0.f};
          const float* arow = am + (bt0 + t) * V + v;
          if (kVec) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(arow));
            a[0] = t4.x; a[1] = t4.y; a[2] = t4.z; a[3] = t4.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = (v + j < V) ? __ldg(arow + j) : 0.f;
          }
          for (int rb = r_lo; rb <= r_hi; rb += kRB) {
            // all loads of up to kRB band slots first, unconditionally (slots outside the band or the chunk read a
            // valid dummy row and are weighted by 0), so that their latencies overlap
            uint2 graw[kRB];
            float l[kRB][4], wgt[kRB];
#pragma unroll
            for (int q = 0; q < kRB; ++q) {
              const int r = rb + q;
              const int64_t m = m_base + r;
              const bool ok = r <= r_hi && m >= row0 && m < row_end;
              wgt[q] = ok ? 1.f : 0.f;
              graw[q] = *reinterpret_cast<const uint2*>(dh + (ok ? (m - row0) * ld : 0) + v);  // ld, v multiples of 4
              const float* lrow = lm + ((int64_t)b * (S + 1) + (ok ? sbt + r : 0)) * V + v;
              if (kVec) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(lrow));
                l[q][0] = t4.x; l[q][1] = t4.y; l[q][2] = t4.z; l[q][3] = t4.w;
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) l[q][j] = (v + j < V) ? __ldg(lrow + j) : 0.f;
              }
            }
#pragma unroll
            for (int q = 0; q < kRB; ++q) {
              const int r = min(rb + q, r_hi);  // a clamped duplicate slot adds 0
              const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(&graw[q]);
              float4* cell = reinterpret_cast<float4*>(mine + (sbt + r - w_lo) * kDjCols);  // sbt + r in [w_lo, w_hi]
              float4 c = *cell;
              float x[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // select, not multiply: the dummy row of a skipped slot may sit in a never-written (padding) block
                x[j] = wgt[q] != 0.f ? __bfloat162float(gb[j]) * act_bwd_fast(a[j] + l[q][j], act) : 0.f;
                dsum[j] += x[j];
              }
              c.x += x[0]; c.y += x[1]; c.z += x[2]; c.w += x[3];
              *cell = c;
            }
          }
          float* drow = d_am + (bt0 + t) * V + v;
          if (kVec) {
            float4 o = make_float4(dsum[0], dsum[1], dsum[2], dsum[3]);
            if (am_accumulate || r_lo > 0) {  // an earlier window already stored this frame's first slots
              const float4 old = *reinterpret_cast<float4*>(drow);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(drow) = o;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (v + j < V) drow[j] = ((am_accumulate || r_lo > 0) ? drow[j] : 0.f) + dsum[j];
          }
        }
      }
      for (int i = 0; i <= w_hi - w_lo; ++i) {
        const float4 c = *reinterpret_cast<const float4*>(mine + i * kDjCols);
        if (c.x == 0.f && c.y == 0.f && c.z == 0.f && c.w == 0.f) continue;
        float* lrow = d_lm + ((int64_t)b * (S + 1) + w_lo + i) * V + v;
        if (kVec) {
          atomicAdd(reinterpret_cast<float4*>(lrow), c);
        } else {
          const float e[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (v + j < V && e[j] != 0.f) atomicAdd(lrow + j, e[j]);
        }
      }
    }
  }
}

template <bool kVec, int kRB>
void launch_djoint_reduce(const JoinerProblem& p, const __nv_bfloat16* dh, int ld, int64_t row0, int64_t rows, int64_t M,
                          bool am_accumulate, float* d_am, float* d_lm, cudaStream_t stream) {
  const dim3 grid((unsigned)((p.T + kDjFrames - 1) / kDjFrames), (unsigned)p.B);
  const size_t smem = (size_t)kDjMaxSpan * kDjCols * sizeof(float);
  auto kern = djoint_reduce_kernel<kVec, kRB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, 128, smem, stream>>>(dh, ld, p.am, p.lm, p.ranges, p.boundary, row0, rows, M, p.T, p.S, p.R, p.V, p.act,
                                    am_accumulate, d_am, d_lm);
}

// C^T accumulate: out[(n + j) * ld + m] += acc[j]     (dW1[i, v] from the (v, i) accumulator)
struct StoreTransposedAtomicEpi {
  static constexpr int kScratchBytes = 0;
  float* out;
  int64_t ld;
  int M, N;  // valid rows (v) and columns (i) of the accumulator
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    if (ctx.m >= M) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < N && acc[j] != 0.f) atomicAdd(out + (int64_t)(n + j) * ld + ctx.m, acc[j]);
  }
};

struct TcDims {
  int64_t M;       // joiner rows
  int Mt;          // row tiles of 128
  int Vp, Ip;      // V, I padded to multiples of 256
  int kbV, kbI;    // K blocks of 64 covering V, Ip
  int n_tiles_v;   // Vp / kBN
  int n_parts_v;   // LSE partials per row
  int64_t chunk;   // rows per backward chunk (multiple of 128)
  bool keep_joint; // keep act(am + lm[ranges]) as a bf16 operand from forward to backward
  bool dead_skip;  // row tiles of padding frames are skipped (their blocks are never written nor read)
};

TcDims tc_dims(int64_t M, int V, int I) {
  TcDims d;
  d.M = M;
  d.Mt = (int)((M + 127) / 128);
  d.Vp = ((V + 255) / 256) * 256;
  d.Ip = ((I + 255) / 256) * 256;
  d.kbV = (V + 63) / 64;
  d.kbI = d.Ip / 64;
  d.n_tiles_v = d.Vp / kBN;
  d.n_parts_v = kLseGroups * d.n_tiles_v;  // one partial per (column tile, epilogue group)
  // device-memory budgets (queried once): an eighth of the memory for the kept J, a 48th (3.7 GB of a B200's 180 GB,
  // at least 1 GiB) for each of the two (chunk rows x Vp) bf16 buffers of the backward pass
  const size_t jp_budget = device_info().total_mem / 8;  // per device: forward and backward of a step agree on the layout
  const size_t budget = jp_budget / 6 > ((size_t)1 << 30) ? jp_budget / 6 : (size_t)1 << 30;
  int64_t rows = (int64_t)(budget / ((size_t)d.Vp * 2));
  if (const char* e = getenv("S2T_B200_CHUNK_ROWS")) rows = atoll(e);  // test hook: force the multi-chunk backward
  rows = (rows / 128) * 128;
  if (rows < 128) rows = 128;
  int64_t all = (int64_t)d.Mt * 128;
  d.chunk = rows < all ? rows : all;
  // Jp is kept while it fits an eighth of the device memory (22 GB of a B200's 180 GB: c5's 13 GB fits); beyond that
  // the hidden and dW1 contractions rebuild it on the fly in their producer warps
  d.keep_joint = (size_t)d.Mt * (d.Vp / 64) * kBlockBytes <= jp_budget && !getenv("S2T_B200_NO_KEEP_JOINT");
  d.dead_skip = d.keep_joint && !getenv("S2T_B200_NO_DEAD_SKIP");  // producer-fed paths touch every row
  return d;
}

// The layout a forward call carved into a workspace, remembered by workspace address: the backward call of the same
// step reads it back instead of re-deriving it from the environment (the row-chunk and keep-J test hooks) and the
// current device, so a change of either between the two calls cannot make them disagree about where things are.
class LayoutMemo {
 public:
  void put(const void* ws, const TcDims& d) {
    std::lock_guard<std::mutex> g(mu_);
    if (!map_.count(ws)) {
      order_.push_back(ws);
      if (order_.size() > kCap) {
        map_.erase(order_.front());
        order_.pop_front();
      }
    }
    map_[ws] = d;
  }
  bool get(const void* ws, int64_t M, TcDims* out) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(ws);
    if (it == map_.end() || it->second.M != M) return false;
    *out = it->second;
    return true;
  }

 private:
  static constexpr size_t kCap = 64;
  std::mutex mu_;
  std::unordered_map<const void*, TcDims> map_;
  std::deque<const void*> order_;
};
LayoutMemo& layout_memo() {
  static LayoutMemo m;
  return m;
}

struct TcWs {
  int* am_row;
  int* lm_row;
  int* row_sym;
  float* part;
  float* sym_logit;
  float* blank_logit;
  uint8_t *W1p, *W2p, *W2Tp, *W1Tp;
  uint8_t* Hp;
  uint8_t* Jp;  // packed act(am + lm[ranges]) (rows m, cols v) or nullptr when it would not fit the budget
  uint8_t *Gp, *DHp;
  __nv_bfloat16* dh;  // (chunk rows, Vp) d loss / d (joint pre-activation) before act'
  uint8_t* tile_live;  // one byte per 128-row tile, or nullptr (J rebuilt on the fly / S2T_B200_NO_DEAD_SKIP)
  int* live_idx;       // the live tiles, ascending
  int* live_prefix;    // live tiles before tile i (Mt + 1 entries)
  size_t bytes;
};

TcWs tc_carve(void* ws, const TcDims& d) {
  TcWs w;
  char* p = (char*)ws;
  auto take = [&](size_t n) {
    char* q = p;
    p += (n + 1023) / 1024 * 1024;
    return q;
  };
  const int ct = (int)(d.chunk / 128);
  w.am_row = (int*)take(d.M * sizeof(int));
  w.lm_row = (int*)take(d.M * sizeof(int));
  w.row_sym = (int*)take(d.M * sizeof(int));
  w.part = (float*)take((size_t)d.M * d.n_parts_v * 2 * sizeof(float));
  w.sym_logit = (float*)take(d.M * sizeof(float));
  w.blank_logit = (float*)take(d.M * sizeof(float));
  w.W1p = (uint8_t*)take((size_t)(d.Ip / 128) * d.kbV * kBlockBytes);
  w.W2Tp = (uint8_t*)take((size_t)(d.Ip / 128) * d.kbV * kBlockBytes);
  w.W2p = (uint8_t*)take((size_t)(d.Vp / 128) * d.kbI * kBlockBytes);
  w.W1Tp = (uint8_t*)take((size_t)(d.Vp / 128) * d.kbI * kBlockBytes);
  w.Hp = (uint8_t*)take((size_t)d.Mt * d.kbI * kBlockBytes);
  w.Jp = d.keep_joint ? (uint8_t*)take((size_t)d.Mt * (d.Vp / 64) * kBlockBytes) : nullptr;
  w.Gp = (uint8_t*)take((size_t)ct * (d.Vp / 64) * kBlockBytes);
  w.DHp = (uint8_t*)take((size_t)ct * d.kbI * kBlockBytes);
  w.dh = (__nv_bfloat16*)take((size_t)d.chunk * d.Vp * sizeof(__nv_bfloat16));
  w.tile_live = (uint8_t*)take((size_t)d.Mt);
  w.live_idx = (int*)take((size_t)d.Mt * sizeof(int));
  w.live_prefix = (int*)take((size_t)(d.Mt + 1) * sizeof(int));
  if (!d.dead_skip) {
    w.tile_live = nullptr;
    w.live_idx = w.live_prefix = nullptr;
  }
  w.bytes = (size_t)(p - (char*)ws);
  return w;
}

// All four operand images of the two weight matrices in one launch (the backward call reuses them: the
// workspace travels unchanged from forward to backward).
int pack_weights(const JoinerProblem& p, const TcDims& d, const TcWs& w, cudaStream_t st) {
  const PackJob jobs[4] = {
      // W1 (I, V): rows i, K v          W2 (V, I): rows v, K i
      {p.W1, p.V, 1, p.I, p.V, d.Ip / 128, d.kbV, w.W1p, 0},
      {p.W2, p.I, 1, p.V, p.I, d.Vp / 128, d.kbI, w.W2p, 0},
      // W2^T: rows i, K v -> element (i, v) = W2[v * I + i]     W1^T: rows v, K i -> W1[i * V + v]
      {p.W2, 1, p.I, p.I, p.V, d.Ip / 128, d.kbV, w.W2Tp, 0},
      {p.W1, 1, p.V, p.V, p.I, d.Vp / 128, d.kbI, w.W1Tp, 0},
  };
  return pack_jobs(jobs, 4, st);
}


// ---- fused forward -----------------------------------------------------------------------------
// hidden = act(am + lm[ranges]) W1^T + b1 and logits = hidden W2^T + b2 -> (lse, px, py) in ONE persistent kernel:
// a CTA owns a 128-row tile of the joiner lattice from the gather to the finished log-probabilities, so that neither
// act(am + lm[ranges]) nor the logits ever exist outside the SM (the hidden rows are written once, for the backward
// pass).  Replaces joint_pack_kernel + the hidden and logits contractions + lse_combine_kernel for inner_dim <= 256.
//
//   warps  8..15  producers: build the A operand of contraction 1, act(am + lm[ranges]) -> bf16, one 64-entry
//                 vocabulary step at a time straight into the swizzled stage image (JointRowProducer)
//   warp   0      bulk copies of W1 (256 x 64 bf16 = 32 KB per step) into the same stage ring
//   warp   2      bulk copies of W2 blocks (128 vocabulary rows x 64 hidden entries = 16 KB) into a second ring
//   warp   1      tcgen05.mma issue.  Contraction 1: 128 x 256 (hidden) accumulator in TMEM columns 0..255, K = V.
//                 Contraction 2: for every 128-column vocabulary tile, 128 x 128 accumulators alternating between
//                 TMEM columns 256..383 and 384..511, K = 256 with the A operand read from the hidden tile that the
//                 epilogue warps parked in shared memory
//   warps  4..7   epilogue (a thread = one lattice row = one TMEM lane): drain the hidden accumulator, add b1, round to
//                 bf16, store the row into the shared-memory K-major image (and the same 16-byte pieces into the
//                 packed hidden operand Hp in HBM); then, per vocabulary tile, the running (max, sum exp) and the
//                 sym / blank gather; lse / px / py are final when the tile's last vocabulary tile has been drained
//
// Shared memory: 2 x 48 KB stage ring 1 (A stage + a W1 k-step) + 64 KB hidden tile + 2 x 16 KB ring 2 (W2 blocks) + the raw
// ring.  What paces the kernel is the weight stream: per 128-row tile every SM pulls W1 and W2 (2 x Vp x 256 x 2 bytes =
// 512 KB at V = 500) through L2, and with all 148 SMs doing so the blocks arrive at ~17 bytes per cycle and SM (~5 TB/s
// over the chip) however many are in flight (role timeline, S2T_TRACE=tc_joiner_fwd_fused: a W2 block every ~1500 cycles
// with 2 blocks in flight, every ~2000 with W1 as half k-steps and three blocks in each weight ring -- Little's law with
// a fixed rate; that variant measured 0.145 ms against 0.134).  Fewer weight bytes per row would need both contractions
// on CTA pairs (cta_group::2: each SM holds half of every weight block).
// role timeline of CTA 0 (S2T_TRACE=tc_joiner_fwd_fused): compiled in with -DS2T_FUSED_TRACE only -- the marks cost the
// kernel ~3 % (registers)
#ifdef S2T_FUSED_TRACE
#define S2T_FUSED_MARK(code, idx) trace_mark(p.trace, code, idx)
#else
#define S2T_FUSED_MARK(code, idx) ((void)0)
#endif
#ifndef S2T_W_PIECES
#define S2T_W_PIECES 1
#endif
constexpr int kWPieces = S2T_W_PIECES;             // bulk copies per 16 KB weight block (4 pieces: 0.136 vs 0.134 ms, no gain)
constexpr int kFS1 = 2;                          // stages of ring 1
constexpr int kFS2 = 2;                          // stages of ring 2
constexpr int kFStage1 = 3 * kBlockBytes;        // A (128 x 64) + B (256 x 64)
constexpr int kFHBytes = 4 * kBlockBytes;        // hidden tile: 128 x 256 bf16
constexpr int kFThreads = 32 * (kCtrlWarps + kEpiWarps + kProdWarps);
// Raw ring: the DISTINCT am / lm rows a tile's 128 lattice rows are built from (a frame's am row serves its R band
// slots, an lm row serves every frame whose band holds that symbol position: ~40 distinct rows against 256 row
// reads), one 64-entry vocabulary step per stage, brought in by bulk copies.  The producers then read them from
// shared memory: what a tile pulls through the L2 -> SM path for its A operand drops from 512 KB to ~80 KB.
// Three stages of 42 rows: what a stage holds is in flight for the ~1.7 us a bulk copy takes under load, so the ring
// depth (not the bandwidth) paces contraction 1 -- with two stages the producers spent 37 % of their time waiting for
// rows (ncu).  42 rows (am rows first, lm rows behind them) cover the ~27 frames + ~13 symbol positions of a tile
// inside one utterance; the tile in ~16 that straddles two utterances takes the direct path.
constexpr int kRawStages = 3;
constexpr int kRawRows = 42;                            // staged rows per step (tiles that need more take the direct path)
constexpr int kRawRowBytes = kBlockK * 4;               // 64 fp32
constexpr int kRawBytes = kRawRows * kRawRowBytes;      // 10.5 KB
constexpr size_t kFSmemBytes = (size_t)kFS1 * kFStage1 + kFHBytes + (size_t)kFS2 * kBlockBytes + (size_t)kRawStages * kRawBytes +
                               (256 + 2 * 128) * sizeof(float) /*b1, two vocabulary tiles of b2*/ + 1024 /*align*/ +
                               512 /*barriers, tile plans*/;

// what the planner warp tells the producers about a tile (four slots, by tile index)
struct TilePlan {
  int fast;      // 1: rows staged in the raw ring; 0: direct global loads (too many distinct rows, or unaligned)
  int am_first;  // first am row (b T + t) of the tile
  int n_am;
  int b0;        // utterance of the tile's first row
  int lo0, n0;   // lm rows (b (S+1) + s) of utterance b0: lo0 .. lo0 + n0 - 1
  int lo1, n1;   // lm rows of utterance b0 + 1
};
static_assert(kFSmemBytes <= 227 * 1024, "fused joiner forward: shared memory");

// Tensor maps of am (B T, V) and lm (B (S+1), V) with boxes of 1, 2, 4, .. 32 rows x 64 columns: the n distinct rows a
// tile needs from a tensor are consecutive rows, fetched as the binary decomposition of n -- ~7 tensor copies per step
// instead of ~40 row copies of 256 bytes (the copy engine's per-instruction cost, not the bytes, paced the raw ring).
constexpr int kRawBoxes = 6;
struct alignas(64) RawMaps {
  CUtensorMap am[kRawBoxes];
  CUtensorMap lm[kRawBoxes];
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// rows x V fp32, row-major; false when the tensor cannot be described (unaligned base, V not a multiple of 4)
bool encode_row_boxes(const float* base, int64_t rows, int V, CUtensorMap* out) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || rows <= 0 || (V & 3) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  for (int i = 0; i < kRawBoxes; ++i) {
    const cuuint64_t dims[2] = {(cuuint64_t)V, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)V * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, 1u << i};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&out[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

struct FusedFwdParams {
  const float* am;
  const float* lm;
  const int* am_row;
  const int* lm_row;
  const int* row_sym;
  const uint8_t* W1p;  // rows i (256), k-blocks over v
  const uint8_t* W2p;  // rows v (Vp), k-blocks over i (4)
  const float* b1;
  const float* b2;
  uint8_t* Hp;         // packed hidden rows (rows m, 4 k-blocks) for the backward pass
  uint8_t* Jp;         // optional by-product: packed act(am + lm[ranges]) for the backward pass
  const int* live_idx;     // live 128-row tiles (or null: all)
  const int* live_prefix;
  const int64_t* boundary;
  float* lse;
  float* px;
  float* py;
  int64_t M;
  int Mt, V, I, kbV, nN, w2_row_blocks, act, blank, T, R;
  float delay_penalty;
  int maps_ok = 0;  // RawMaps describe am / lm (set by the launcher): the raw ring is fed by tensor copies
  int raw_stages = kRawStages, raw_rows = kRawRows;  // the ring's 126 row slots as 3 x 42 (pruned bands) or 2 x 63 (wide bands)
  unsigned long long* trace = nullptr;  // S2T_TRACE=tc_joiner_fwd_fused: role timeline of CTA 0
};

template <int kAct>
__global__ void __launch_bounds__(kFThreads, 1) joiner_fwd_fused_kernel(const FusedFwdParams p,
                                                                        const __grid_constant__ RawMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* ring1 = smem;
  uint8_t* H = ring1 + kFS1 * kFStage1;
  uint8_t* ring2 = H + kFHBytes;
  uint8_t* raw = ring2 + kFS2 * kBlockBytes;
  float* sb1 = reinterpret_cast<float*>(raw + kRawStages * kRawBytes);  // b1, zero beyond I
  float* sb2 = sb1 + 256;                                               // b2 of the vocabulary tile in flight, x 2
  uint64_t* full1 = reinterpret_cast<uint64_t*>(sb2 + 2 * 128);
  uint64_t* empty1 = full1 + kFS1;
  uint64_t* full2 = empty1 + kFS1;
  uint64_t* empty2 = full2 + kFS2;
  uint64_t* acc1_full = empty2 + kFS2;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* h_full = acc1_empty + 1;
  uint64_t* h_empty = h_full + 1;
  uint64_t* acc2_full = h_empty + 1;   // [2]
  uint64_t* acc2_empty = acc2_full + 2;  // [2]
  uint64_t* raw_full = acc2_empty + 2;   // [kRawStages]
  uint64_t* raw_empty = raw_full + kRawStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kRawStages);
  TilePlan* plans = reinterpret_cast<TilePlan*>(tmem_slot + 4);  // [4]: the planner runs at most kRawStages steps ahead

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFS1; ++s) {
      mbar_init(&full1[s], 1 + kProdWarps);
      mbar_init(&empty1[s], 1);
    }
    for (int s = 0; s < kFS2; ++s) {
      mbar_init(&full2[s], 1);
      mbar_init(&empty2[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, kEpiWarps);
    mbar_init(h_full, kEpiWarps);
    mbar_init(h_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], kEpiWarps);
    }
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kProdWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // programmatic dependent launch: barrier init and the TMEM allocation overlap the tail of the previous kernel; no
  // global memory is read before it has completed
  griddep_wait();
  griddep_launch_dependents();
  const bool listed = p.live_idx != nullptr;
  const int n_tiles = listed ? p.live_prefix[p.Mt] : p.Mt;
  auto tile_of = [&](int j) { return listed ? p.live_idx[j] : j; };
  // the bias values live in shared memory: the epilogue reads 32 of them per accumulator chunk, and with 224 KB of
  // shared memory in use the L1 that would otherwise serve those loads is a few KB
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sb1[i] = i < p.I ? __ldg(p.b1 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int raw_stages = p.raw_stages, raw_bytes = p.raw_rows * kRawRowBytes;

  if (warp == 0) {
    // ---- W1 steps into ring 1 ----
    if (lane == 0) {
      uint32_t g = 0;
      for (int j = blockIdx.x; j < n_tiles; j += gridDim.x) {
        for (int ks = 0; ks < p.kbV; ++ks, ++g) {
          const int s = g % kFS1;
          mbar_wait(&empty1[s], ((g / kFS1) & 1) ^ 1);
          mbar_arrive_expect_tx(&full1[s], 2 * kBlockBytes);
          {
            uint8_t* dst = ring1 + s * kFStage1 + kBlockBytes;
            const uint8_t* src = p.W1p + (size_t)ks * 2 * kBlockBytes;
#pragma unroll
            for (int q = 0; q < kWPieces * 2; ++q)
              bulk_copy_g2s(dst + q * (kBlockBytes / kWPieces), src + q * (kBlockBytes / kWPieces), kBlockBytes / kWPieces, &full1[s]);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---- W2 blocks into ring 2 ----
    if (lane == 0) {
      uint32_t g = 0;
      for (int j = blockIdx.x; j < n_tiles; j += gridDim.x) {
        for (int n = 0; n < p.nN; ++n) {
          for (int kb = 0; kb < 4; ++kb, ++g) {
            const int s = g % kFS2;
            mbar_wait(&empty2[s], ((g / kFS2) & 1) ^ 1);
            mbar_arrive_expect_tx(&full2[s], kBlockBytes);
            {
              uint8_t* dst = ring2 + s * kBlockBytes;
              const uint8_t* src = p.W2p + packed_block_index(n, kb, p.w2_row_blocks) * kBlockBytes;
#pragma unroll
              for (int q = 0; q < kWPieces; ++q)
                bulk_copy_g2s(dst + q * (kBlockBytes / kWPieces), src + q * (kBlockBytes / kWPieces), kBlockBytes / kWPieces, &full2[s]);
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ---- planner: which distinct am / lm rows a tile needs, and their bulk copies into the raw ring ----
    const bool aligned = (p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.am) | reinterpret_cast<uintptr_t>(p.lm)) & 15) == 0;
    uint32_t g = 0, lt = 0;
    // the row indices of a tile (am_row / lm_row of its 128 rows: four per lane) are fetched one tile ahead, so that
    // their round trip through L2 runs under the previous tile's copies instead of in front of this tile's first one
    int ar_q[4], lr_q[4], a_first = 0;
    auto fetch_rows = [&](int jj) {
      const int64_t mm0 = (int64_t)tile_of(jj) * 128;
      a_first = __ldg(p.am_row + mm0);  // row m0 < M: the tile is live
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t m = mm0 + lane + 32 * q;
        ar_q[q] = m < p.M ? __ldg(p.am_row + m) : -1;
        lr_q[q] = m < p.M ? __ldg(p.lm_row + m) : 0;
      }
    };
    if ((int)blockIdx.x < n_tiles) fetch_rows(blockIdx.x);
    for (int j = blockIdx.x; j < n_tiles; j += gridDim.x, ++lt) {
      // the stage of the tile's first step must be free before a plan slot is overwritten (four slots, the planner
      // is at most kRawStages steps ahead of the producers)
      mbar_wait(&raw_empty[g % raw_stages], ((g / raw_stages) & 1) ^ 1);
      int am_lo = INT_MAX, am_hi = -1, lo0 = INT_MAX, hi0 = -1, lo1 = INT_MAX, hi1 = -1, bad = 0;
      const int b0 = a_first / p.T;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ar_q[q] >= 0) {
          const int ar = ar_q[q], lr = lr_q[q];
          am_lo = min(am_lo, ar);
          am_hi = max(am_hi, ar);
          const int seg = ar / p.T - b0;
          if (seg == 0) {
            lo0 = min(lo0, lr);
            hi0 = max(hi0, lr);
          } else if (seg == 1) {
            lo1 = min(lo1, lr);
            hi1 = max(hi1, lr);
          } else {
            bad = 1;
          }
        }
      }
      if (j + (int)gridDim.x < n_tiles) fetch_rows(j + gridDim.x);  // next tile's indices: in flight during this tile's steps
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        am_lo = min(am_lo, __shfl_xor_sync(0xffffffffu, am_lo, o));
        am_hi = max(am_hi, __shfl_xor_sync(0xffffffffu, am_hi, o));
        lo0 = min(lo0, __shfl_xor_sync(0xffffffffu, lo0, o));
        hi0 = max(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
        lo1 = min(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
        hi1 = max(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
      }
      const int n_am = am_hi - am_lo + 1, n0 = hi0 >= lo0 ? hi0 - lo0 + 1 : 0, n1 = hi1 >= lo1 ? hi1 - lo1 + 1 : 0;
      const bool fast = aligned && p.maps_ok && !bad && n_am + n0 + n1 <= p.raw_rows;
      if (lane == 0) plans[lt & 3] = TilePlan{fast ? 1 : 0, am_lo, n_am, b0, lo0, n0, lo1, n1};
      __syncwarp();
      const int n_rows = n_am + n0 + n1;
      // this lane's tensor copy of every step: lanes 0..5 the set bits of n_am (32, 16, .. 1 rows), lanes 6..11 those of
      // n0, lanes 12..17 those of n1; a box of 2^bit rows sits behind the boxes of the higher bits of its segment
      int my_h = 0, my_bit = 0, my_row = 0, my_slot = 0;
      bool my_lm = false;
      if (fast && lane < 3 * kRawBoxes) {
        const int seg = lane / kRawBoxes;
        my_bit = kRawBoxes - 1 - lane % kRawBoxes;
        const int n = seg == 0 ? n_am : (seg == 1 ? n0 : n1);
        const int first = seg == 0 ? am_lo : (seg == 1 ? lo0 : lo1);
        const int slot0 = seg == 0 ? 0 : (seg == 1 ? n_am : n_am + n0);
        if (n & (1 << my_bit)) {
          const int before = (n >> (my_bit + 1)) << (my_bit + 1);  // rows covered by the higher bits
          my_h = 1 << my_bit;
          my_row = first + before;
          my_slot = slot0 + before;
          my_lm = seg != 0;
        }
      }
      for (int ks = 0; ks < p.kbV; ++ks, ++g) {
        const int s = g % raw_stages;
        if (ks > 0) mbar_wait(&raw_empty[s], ((g / raw_stages) & 1) ^ 1);
        if (!fast) {
          if (lane == 0) mbar_arrive(&raw_full[s]);
          continue;
        }
        // a box arrives whole (columns beyond V as zeros): 256 bytes per row whatever the step
        if (lane == 0) mbar_arrive_expect_tx(&raw_full[s], (uint32_t)kRawRowBytes * (uint32_t)n_rows);
        __syncwarp();
        uint8_t* dst0 = raw + s * raw_bytes;
        if (lane == 0) S2T_FUSED_MARK(7, (int)g);
        if (my_h > 0)
          tma_load_2d(dst0 + my_slot * kRawRowBytes, my_lm ? (const void*)&maps.lm[my_bit] : (const void*)&maps.am[my_bit],
                      ks * kBlockK, my_row, &raw_full[s]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue ----
    // One thread serves both contractions.  Contraction 1 of tile i + 1 (fed by the producers) and contraction 2 of
    // tile i (paced by the log-sum-exp epilogue) use different accumulators, stage rings and hand-over barriers, so
    // they are interleaved: the thread polls both pipelines and issues whichever has a stage ready -- blocking on
    // one would serialise "build the hidden rows" and "reduce the logits", each of which keeps other warps busy.
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_bf16(128, 256), idesc2 = umma_idesc_bf16(128, 128);
      const int n_mine = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      const uint32_t sh = smem_u32(H);
      uint32_t g1 = 0, g2 = 0, nt = 0;
      int t1 = 0, k1 = 0;           // contraction 1: tile (in this CTA's sequence), vocabulary step
      int t2 = 0, n2 = 0, kb2 = 0;  // contraction 2: tile, vocabulary tile, hidden k-block
      bool acc1_ready = false, h_ready = false, acc2_ready = false;
      uint32_t idle = 0;
      while (t2 < n_mine) {
        bool progressed = false;
        if (t2 < t1) {  // the hidden rows of tile t2 are (being) produced: contraction 2
          if (!h_ready) h_ready = mbar_test(h_full, t2 & 1);
          if (h_ready) {
            const uint32_t buf = nt & 1;
            if (!acc2_ready) acc2_ready = mbar_test(&acc2_empty[buf], ((nt >> 1) & 1) ^ 1);
            const int s = g2 % kFS2;
            if (acc2_ready && mbar_test(&full2[s], (g2 / kFS2) & 1)) {
              tc_fence_after();
              const uint32_t acc = tmem_base + 256 + buf * 128;
              const uint32_t sa = sh + kb2 * kBlockBytes, sb = smem_u32(ring2 + s * kBlockBytes);
#pragma unroll
              for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4)
                umma_bf16(acc, umma_smem_desc(sa + k4 * kUmmaK * 2), umma_smem_desc(sb + k4 * kUmmaK * 2), idesc2,
                          kb2 > 0 || k4 > 0);
              umma_commit(&empty2[s]);
              ++g2;
              progressed = true;
              if (++kb2 == 4) {
                umma_commit(&acc2_full[buf]);
                S2T_FUSED_MARK(10, (int)nt);
                kb2 = 0;
                ++nt;
                acc2_ready = false;
                if (++n2 == p.nN) {
                  umma_commit(h_empty);  // every MMA that reads this tile's hidden rows has completed
                  n2 = 0;
                  ++t2;
                  h_ready = false;
                }
              }
            }
          }
        }
        if (t1 < n_mine) {  // contraction 1 of tile t1 -> TMEM columns 0..255 (drained by the epilogue of tile t1 - 1)
          if (!acc1_ready) acc1_ready = mbar_test(acc1_empty, (t1 & 1) ^ 1);
          const int s = g1 % kFS1;
          if (acc1_ready && mbar_test(&full1[s], (g1 / kFS1) & 1)) {
            tc_fence_after();
            const uint32_t sa = smem_u32(ring1 + s * kFStage1), sb = sa + kBlockBytes;
#pragma unroll
            for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4)
              umma_bf16(tmem_base, umma_smem_desc(sa + k4 * kUmmaK * 2), umma_smem_desc(sb + k4 * kUmmaK * 2), idesc1,
                        k1 > 0 || k4 > 0);
            umma_commit(&empty1[s]);
            S2T_FUSED_MARK(3, (int)g1);
            ++g1;
            progressed = true;
            if (++k1 == p.kbV) {
              umma_commit(acc1_full);
              S2T_FUSED_MARK(4, t1);
              k1 = 0;
              ++t1;
              acc1_ready = false;
            }
          }
        }
        if (progressed) {
          idle = 0;
        } else if (++idle > (1u << 28)) {
          __trap();  // a protocol bug must not hang the GPU
        }
      }
    }
  } else if (warp >= kCtrlWarps && warp < kCtrlWarps + kEpiWarps) {
    // ---- epilogue ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t nt = 0, lt = 0;
    for (int j = blockIdx.x; j < n_tiles; j += gridDim.x, ++lt) {
      const int tile = tile_of(j);
      const int64_t m = (int64_t)tile * 128 + r;
      const bool live = m < p.M;
      // hidden rows: TMEM -> + b1 -> bf16 -> shared-memory operand image (+ Hp)
      mbar_wait(acc1_full, lt & 1);
      if (warp == kCtrlWarps && lane == 0) S2T_FUSED_MARK(5, (int)lt);
      tc_fence_after();
      mbar_wait(h_empty, (lt & 1) ^ 1);
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        float v[32];
        tmem_ld_32x32(lane_addr + cc * 32, v);
        const int n = cc * 32;
        float x[32];
#pragma unroll
        for (int q = 0; q < 32; q += 4) {  // columns beyond I hold exact zeros (zero rows of W1, zero bias)
          const float4 bq = *reinterpret_cast<const float4*>(sb1 + n + q);
          x[q] = live ? v[q] + bq.x : 0.f;
          x[q + 1] = live ? v[q + 1] + bq.y : 0.f;
          x[q + 2] = live ? v[q + 2] + bq.z : 0.f;
          x[q + 3] = live ? v[q + 3] + bq.w : 0.f;
        }
        uint4 mine[4];
        pack_row32_bf16(x, mine);
        const int kb = n >> 6, c0 = (n & 63) >> 3;
        uint8_t* hs = H + kb * kBlockBytes;
        uint8_t* hg = p.Hp + packed_block_index(tile, kb, p.Mt) * kBlockBytes;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = block_chunk_offset(r, c0 + q);
          *reinterpret_cast<uint4*>(hs + off) = mine[q];
          *reinterpret_cast<uint4*>(hg + off) = mine[q];  // the four pieces fill one 64-byte half of the row's block row
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc1_empty);
        mbar_arrive(h_full);
        if (warp == kCtrlWarps && lane == 0) S2T_FUSED_MARK(6, (int)lt);
      }
      // logits: running (max, sum exp) over the vocabulary tiles + the sym / blank gather
      float mx = kNegInf, sum = 0.f, sym_logit = 0.f, blank_logit = 0.f;
      const int csym = live ? __ldg(p.row_sym + m) : -1;
      for (int nn = 0; nn < p.nN; ++nn, ++nt) {
        const uint32_t buf = nt & 1;
        // this vocabulary tile's 128 bias values (-inf beyond V: those columns drop out of max and sum)
        float* b2s = sb2 + buf * 128;
        {
          const int col = nn * 128 + r;
          b2s[r] = col < p.V ? __ldg(p.b2 + col) : kNegInf;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps; also orders the reuse two tiles on
        mbar_wait(&acc2_full[buf], (nt >> 1) & 1);
        if (warp == kCtrlWarps && lane == 0) S2T_FUSED_MARK(11, (int)nt);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          float v[32];
          tmem_ld_32x32(lane_addr + 256 + buf * 128 + cc * 32, v);
          const int n = nn * 128 + cc * 32;
          if (!live || n >= p.V) continue;
          float x[32];
          float cm = kNegInf;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 bq = *reinterpret_cast<const float4*>(b2s + cc * 32 + q);
            x[q] = v[q] + bq.x;
            x[q + 1] = v[q + 1] + bq.y;
            x[q + 2] = v[q + 2] + bq.z;
            x[q + 3] = v[q + 3] + bq.w;
            cm = fmaxf(cm, fmaxf(fmaxf(x[q], x[q + 1]), fmaxf(x[q + 2], x[q + 3])));
          }
          if (cm > mx) {
            sum *= ex2_approx((mx - cm) * kLog2e);
            mx = cm;
          }
          const float nm = -mx * kLog2e;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            s0 += ex2_approx(fmaf(x[q], kLog2e, nm));
            s1 += ex2_approx(fmaf(x[q + 1], kLog2e, nm));
          }
          sum += s0 + s1;
          if (csym >= n && csym < n + 32) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (n + q == csym) sym_logit = x[q];
          }
          if (p.blank >= n && p.blank < n + 32) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (n + q == p.blank) blank_logit = x[q];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[buf]);
        if (warp == kCtrlWarps && lane == 0) S2T_FUSED_MARK(12, (int)nt);
      }
      if (live) {
        const int64_t bt = m / p.R;
        const int b = (int)(bt / p.T), t = (int)(bt % p.T);
        const int Tb = p.boundary ? (int)p.boundary[4 * b + 3] : p.T;
        if (p.boundary && t >= min(max(Tb, 0), p.T)) {  // padding frame: the lattice ignores these entries
          p.lse[m] = 0.f;
          p.px[m] = 0.f;
          p.py[m] = 0.f;
        } else {
          const float l = mx + logf(sum);
          float xv = sym_logit - l;
          if (p.delay_penalty != 0.f) xv += p.delay_penalty * (0.5f * (float)(Tb - 1) - (float)t);
          p.lse[m] = l;
          p.px[m] = xv;
          p.py[m] = blank_logit - l;
        }
      }
    }
  } else if (warp >= kCtrlWarps + kEpiWarps) {
    // ---- producers: act(am + lm[ranges]) -> ring 1 ----
    // Thread (warp w, lane l) builds 8 consecutive lattice rows, (2 w + l / 16) * 8 + i, and the four vocabulary
    // entries 4 (l % 16) .. + 3 of every step: sixteen lanes cover the 256 bytes a row contributes to a step.
    const JointRowProducer direct{p.am, p.lm, p.am_row, p.lm_row, p.M, p.V, p.act, p.Jp, p.Mt};
    const int pw = warp - kCtrlWarps - kEpiWarps;
    const int c = lane & 15, rbase = (pw * 2 + (lane >> 4)) * 8;
    uint32_t g = 0, graw = 0, lt = 0;
    // this thread's eight (am row, lm row) pairs, fetched one tile ahead like the planner's
    int ar_n[8], lr_n[8];
    auto fetch_rows = [&](int jj) {
      const int64_t mm0 = (int64_t)tile_of(jj) * 128 + rbase;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ar_n[i] = mm0 + i < p.M ? __ldg(p.am_row + mm0 + i) : -1;
        lr_n[i] = mm0 + i < p.M ? __ldg(p.lm_row + mm0 + i) : 0;
      }
    };
    if ((int)blockIdx.x < n_tiles) fetch_rows(blockIdx.x);
    for (int j = blockIdx.x; j < n_tiles; j += gridDim.x, ++lt) {
      const int tile = tile_of(j);
      mbar_wait(&raw_full[graw % raw_stages], (graw / raw_stages) & 1);  // also publishes the tile's plan
      const TilePlan plan = plans[lt & 3];
      if (!plan.fast) {
        if (j + (int)gridDim.x < n_tiles) fetch_rows(j + gridDim.x);
        // the raw ring carries nothing for this tile: hand its stages straight back, then load directly
        for (int ks = 0; ks < p.kbV; ++ks, ++graw) {
          if (ks > 0) mbar_wait(&raw_full[graw % raw_stages], (graw / raw_stages) & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[graw % raw_stages]);
        }
        ProdCtx pc;
        pc.m_tile = tile;
        pc.n_tile = 0;
        pc.batch = 0;
        pc.ks0 = 0;
        pc.n_it = p.kbV;
        pc.valid = true;
        pc.t = pw * 32 + lane;
        pc.smem = ring1;
        pc.stage_bytes = kFStage1;
        pc.stages = kFS1;
        pc.it0 = g;
        pc.full = full1;
        pc.empty = empty1;
        direct.run(pc);
        g += p.kbV;
        continue;
      }
      int a_off[8], l_off[8];  // byte offsets of the row's staged am / lm row inside a raw stage; -1: dead row
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a_off[i] = -1;
        l_off[i] = 0;
        if (ar_n[i] >= 0) {
          const int ar = ar_n[i], lr = lr_n[i];
          a_off[i] = (ar - plan.am_first) * kRawRowBytes + c * 16;
          const int slot = (ar / p.T == plan.b0) ? lr - plan.lo0 : plan.n0 + lr - plan.lo1;
          l_off[i] = (plan.n_am + slot) * kRawRowBytes + c * 16;
        }
      }
      if (j + (int)gridDim.x < n_tiles) fetch_rows(j + gridDim.x);  // in flight during this tile's steps
      for (int ks = 0; ks < p.kbV; ++ks, ++g, ++graw) {
        const int sr = graw % raw_stages;
        if (ks > 0) mbar_wait(&raw_full[sr], (graw / raw_stages) & 1);
        if (pw == 0 && lane == 0) S2T_FUSED_MARK(8, (int)graw);
        const uint8_t* rs = raw + sr * raw_bytes;
        const int v = ks * kBlockK + c * 4;
        uint2 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool live = a_off[i] >= 0;
          const float4 a = *reinterpret_cast<const float4*>(rs + (live ? a_off[i] : 0));
          const float4 l = *reinterpret_cast<const float4*>(rs + (live ? l_off[i] : 0));
          // Dead rows and the vocabulary padding must come out as exact zeros (the stage may hold stale bits there):
          // the INPUT is selected to 0 and act(0) = 0, so that the activation itself is straight-line code -- a
          // conditional around it compiles to one divergence region per element, which exposes the latency of every
          // single MUFU instead of pipelining the 32 of a step.
          const float s0 = (live && v + 0 < p.V) ? a.x + l.x : 0.f;
          const float s1 = (live && v + 1 < p.V) ? a.y + l.y : 0.f;
          const float s2 = (live && v + 2 < p.V) ? a.z + l.z : 0.f;
          const float s3 = (live && v + 3 < p.V) ? a.w + l.w : 0.f;
          if (kAct == kRelu)
            o[i] = make_uint2(pack_bf16x2(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), pack_bf16x2(fmaxf(s2, 0.f), fmaxf(s3, 0.f)));
          else
            o[i] = make_uint2(pack_bf16x2(tanh_fast(s0), tanh_fast(s1)), pack_bf16x2(tanh_fast(s2), tanh_fast(s3)));
        }
        // the raw stage is consumed (its values sit in registers): the planner may refill it
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[sr]);
        const int s1 = g % kFS1;
        mbar_wait(&empty1[s1], ((g / kFS1) & 1) ^ 1);
        uint8_t* dst = ring1 + s1 * kFStage1;
        uint8_t* jdst = p.Jp ? p.Jp + packed_block_index(tile, ks, p.Mt) * kBlockBytes : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rbase + i;
          const uint32_t off = (uint32_t)(r * 128 + ((((c >> 1) ^ (r & 7)) & 7) << 4) + (c & 1) * 8);
          *reinterpret_cast<uint2*>(dst + off) = o[i];
          if (jdst) *reinterpret_cast<uint2*>(jdst + off) = o[i];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full1[s1]);
        if (pw == 0 && lane == 0) S2T_FUSED_MARK(9, (int)g);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int launch_joiner_fwd_fused(FusedFwdParams p, int64_t am_rows, int64_t lm_rows, cudaStream_t stream) {
  RawMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (p.R > 8) {  // wide bands (the unpruned lattice: R = S + 1 rows per frame share one am row and need R lm rows)
    p.raw_stages = 2;
    p.raw_rows = kRawStages * kRawRows / 2;
  }
  p.maps_ok = encode_row_boxes(p.am, am_rows, p.V, maps.am) && encode_row_boxes(p.lm, lm_rows, p.V, maps.lm) ? 1 : 0;
  static bool configured[2][kMaxDevices] = {};
  const bool relu = p.act == kRelu;
  if (int rc = relu ? configure_smem_once(joiner_fwd_fused_kernel<kRelu>, kFSmemBytes, configured[0], "tc_joiner_fwd_fused")
                    : configure_smem_once(joiner_fwd_fused_kernel<kTanh>, kFSmemBytes, configured[1], "tc_joiner_fwd_fused"))
    return rc;
  const int sms = device_info().sms;
  // upper bound of the live tiles (the exact count lives on the device): CTAs beyond it find no tile and leave
  int grid = p.Mt < sms ? p.Mt : sms;
  if (const char* e = getenv("S2T_FUSED_GRID")) grid = atoi(e) < grid ? atoi(e) : grid;  // experiment: fewer CTAs
  tc::MnDebug trace_mn;
  tc::TraceScope trace_scope("tc_joiner_fwd_fused", stream, trace_mn);
  p.trace = trace_mn.trace;
  ProfScope prof("tc_joiner_fwd_fused", stream);
  const cudaError_t e =
      relu ? tc::launch_pdl(joiner_fwd_fused_kernel<kRelu>, (unsigned)grid, kFThreads, kFSmemBytes, stream, 1, p, maps)
           : tc::launch_pdl(joiner_fwd_fused_kernel<kTanh>, (unsigned)grid, kFThreads, kFSmemBytes, stream, 1, p, maps);
  if (e != cudaSuccess) {
    set_error("tc_joiner_fwd_fused: launch: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("tc_joiner_fwd_fused");
}

bool joiner_fused_fwd_ok(const TcDims& d) {
  return getenv("S2T_B200_NO_FUSED_FWD") == nullptr && d.Ip == 256;  // read per call: a test hook toggles it
}

}  // namespace

size_t joiner_tc_workspace_bytes(int64_t M, int V, int I) {
  TcDims d = tc_dims(M, V, I > 0 ? I : 256);
  TcWs w = tc_carve(nullptr, d);
  return w.bytes + 4096;
}

int joiner_tc_forward(const JoinerProblem& p, void* workspace, float* lse, float* px, float* py,
                      cudaStream_t stream) {
  S2T_REQUIRE(p.I > 0, "bf16 tensor-core joiner needs the out-projection (use_out_project=True); "
                       "the projection-free joiner has no contraction and runs in fp32 mode");
  const int64_t M = (int64_t)p.B * p.T * p.R;
  if (M == 0) return 0;
  TcDims d = tc_dims(M, p.V, p.I);
  layout_memo().put(workspace, d);
  TcWs w = tc_carve(workspace, d);
  ForkJoin fj(stream);
  if (w.tile_live) {
    // one small block, needed first by the J pre-pass: runs beside the row metadata and the weight packs
    tc_live_tiles_kernel<<<1, 1024, 0, fj.side(0)>>>(p.boundary, M, p.T, p.R, d.Mt, w.tile_live, w.live_idx, w.live_prefix);
  }
  {
    ProfScope prof("tc_row_meta_kernel", stream);
    tc_row_meta_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(p.ranges, p.sym, M, p.T, p.R, p.S, p.blank,
                                                                       w.am_row, w.lm_row, w.row_sym);
  }
  if (int rc = check_launch("tc_row_meta_kernel")) return rc;
  if (int rc = pack_weights(p, d, w, stream)) return rc;
  fj.join();
  MnDebug live;
  live.live_idx = w.live_idx;
  live.live_prefix = w.live_prefix;
  if (joiner_fused_fwd_ok(d)) {
    // one kernel from the gather to (lse, px, py); tiles of padding frames are never computed: their entries are zeros
    if (w.live_idx) {
      if (px == lse + M && py == px + M) {  // one buffer (functional.py allocates them so): one memset node
        zero_async(lse, (size_t)3 * M * sizeof(float), stream);
      } else {
        cudaMemsetAsync(lse, 0, (size_t)M * sizeof(float), stream);
        cudaMemsetAsync(px, 0, (size_t)M * sizeof(float), stream);
        cudaMemsetAsync(py, 0, (size_t)M * sizeof(float), stream);
      }
    }
    FusedFwdParams fp{p.am, p.lm, w.am_row, w.lm_row, w.row_sym, w.W1p, w.W2p, p.b1, p.b2, w.Hp, w.Jp, w.live_idx,
                      w.live_prefix, p.boundary, lse, px, py, M, d.Mt, p.V, p.I, d.kbV, d.Vp / 128, d.Vp / 128, p.act,
                      p.blank, p.T, p.R, p.delay_penalty};
    return launch_joiner_fwd_fused(fp, (int64_t)p.B * p.T, (int64_t)p.B * (p.S + 1), stream);
  }
  // hidden: M x Ip, K = V.  With J kept for the backward pass it is written once by a fully parallel kernel and the
  // contraction streams it like any packed operand; otherwise the producer warps build it on the fly.
  {
    HiddenEpi ep{p.b1, p.I, M, w.Hp, d.Mt};
    if (w.Jp) {
      {
        ProfScope prof("joint_pack_kernel", stream);
        const unsigned grid = (unsigned)(d.Mt * (128 / kJpRows));
        if (p.act == kRelu)
          joint_pack_kernel<kRelu><<<grid, 256, 0, stream>>>(p.am, p.lm, w.am_row, w.lm_row, M, p.V, d.Mt, d.Vp / 64, w.Jp,
                                                             w.tile_live);
        else
          joint_pack_kernel<kTanh><<<grid, 256, 0, stream>>>(p.am, p.lm, w.am_row, w.lm_row, M, p.V, d.Mt, d.Vp / 64, w.Jp,
                                                             w.tile_live);
      }
      if (int rc = check_launch("joint_pack_kernel")) return rc;
      BulkA a{w.Jp, d.Mt};
      if (int rc = launch_gemm_stream<kBN, kNStages, false, 0>(a, w.W1p, d.Ip / 128, d.Mt, d.Ip / kBN, d.kbV, 1, ep, stream,
                                                               "tc_joiner_hidden_gemm", live))
        return rc;
    } else {
      JointRowProducer a{p.am, p.lm, w.am_row, w.lm_row, M, p.V, p.act, nullptr, d.Mt};
      if (int rc = launch_gemm_stream<kBN, kNStages, false, 0, kPair>(a, w.W1p, d.Ip / 128, d.Mt, d.Ip / kBN, d.kbV, 1, ep, stream,
                                                               "tc_joiner_hidden_gemm"))
        return rc;
    }
  }
  // logits -> lse partials: M x Vp, K = Ip
  {
    BulkA a{w.Hp, d.Mt};
    LseEpi ep{p.b2, w.row_sym, p.V, p.blank, d.n_parts_v, M, w.part, w.sym_logit, w.blank_logit};
    if (int rc = launch_gemm_bstationary<kBN, kNStagesRes, kResSteps>(a, w.W2p, d.Vp / 128, d.Mt, d.n_tiles_v, d.kbI, ep, stream,
                                                                     "tc_joiner_logits_lse_gemm", live))
      return rc;
  }
  {
    ProfScope prof("lse_combine_kernel", stream);
    lse_combine_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(w.part, w.sym_logit, w.blank_logit, p.boundary,
                                                                       M, d.n_parts_v, p.T, p.R, p.delay_penalty, lse,
                                                                       px, py);
  }
  return check_launch("lse_combine_kernel");
}

// Gradients are ACCUMULATED into d_lm, dW1, db1, dW2, db2 (the caller zero-fills); d_am is fully written here.
int joiner_tc_backward(const JoinerProblem& p, void* workspace, const float* lse, const float* occ_px,
                       const float* occ_py, const float* coef, float clamp, float* d_am, float* d_lm, float* dW1,
                       float* db1, float* dW2, float* db2, cudaStream_t stream) {
  const int64_t M = (int64_t)p.B * p.T * p.R;
  if (M == 0) return 0;
  TcDims d = tc_dims(M, p.V, p.I);
  {
    TcDims fwd;  // what the forward call of this workspace used
    if (layout_memo().get(workspace, M, &fwd) && fwd.Vp == d.Vp && fwd.Ip == d.Ip) d = fwd;
  }
  TcWs w = tc_carve(workspace, d);
  const int sms = device_info().sms;
  // d_am: with one row chunk the dJoint reduction writes every row itself (plain first stores, zeros for padding
  // frames); with several chunks a frame can straddle two launches, which then accumulate into a zero-filled buffer
  if (d.chunk < (int64_t)d.Mt * 128) cudaMemsetAsync(d_am, 0, (size_t)p.B * p.T * p.V * sizeof(float), stream);
  for (int64_t row0 = 0; row0 < (int64_t)d.Mt * 128; row0 += d.chunk) {
    const int64_t rows_pad = ((int64_t)d.Mt * 128 - row0 < d.chunk) ? ((int64_t)d.Mt * 128 - row0) : d.chunk;
    const int ct = (int)(rows_pad / 128);       // row tiles of this chunk
    const int kbM = (int)(rows_pad / 64);       // K blocks when rows are the contraction index
    const int tile0 = (int)(row0 / 128);
    MnDebug live;  // live row tiles of this chunk
    live.live_idx = w.live_idx;
    live.live_prefix = w.live_prefix;
    live.live_off = tile0;
    // G = d loss / d logits of the chunk (+ db2)
    {
      BulkA a{w.Hp + (size_t)tile0 * kBlockBytes, d.Mt};  // block(rb, kb) = kb * Mt + rb: shift rb by tile0
      GradEpi ep{p.b2, w.row_sym, lse, occ_px, occ_py, coef, row0, M, p.T * p.R, p.V, p.blank, clamp, w.Gp, ct, db2};
      // (a 128-column resident tile with a 7-stage A ring instead of 256 columns / 3 stages: 0.071 -> 0.083 ms)
      if (int rc = launch_gemm_bstationary<kBN, kNStagesRes, kResSteps>(a, w.W2p, d.Vp / 128, ct, d.n_tiles_v, d.kbI, ep, stream,
                                                                       "tc_joiner_grad_logits_gemm", live))
        return rc;
    }
    // Three independent chains follow G: dW2 (needs G and the hidden activations), and after dhidden, dW1 and dh -> dJ.
    // The weight-gradient contractions run on side streams so that their CTAs fill the tails of the main chain.
    ForkJoin fj(stream);
    cudaStream_t s_dw2 = fj.side(0);  // forked here: after G, before dhidden
    // dhidden = G W2: rows m, N = Ip, K = V
    {
      BulkA a{w.Gp, ct};
      DHiddenEpi ep{p.I, w.DHp, ct, db1};
      // (W2^T resident per 128-column tile and G streamed twice instead: 0.0455 -> 0.060 ms at c3)
      // (as a CTA pair over the live-tile list -- half of W2^T per SM: 0.0457 vs 0.0455 ms, no gain)
      if (int rc = launch_gemm_stream<kBN, kNStages, false, 0>(a, w.W2Tp, d.Ip / 128, ct, d.Ip / kBN, d.kbV, 1, ep, stream,
                                                               "tc_joiner_dhidden_gemm", live))
        return rc;
    }
    const int splits = max(1, min(kbM, sms / max(1, (d.Vp / 128) * (d.Ip / kBN))));
    // dW2[v, i] += sum_m G[m, v] hidden[m, i]: both operands MN-major (contraction over the rows m)
    {
      BulkA a{w.Gp, ct};
      StoreRowMajorEpi ep{dW2, p.I, p.V, p.I, true};
      if (int rc = launch_gemm_stream<kBN, kNStages, true, 0>(a, w.Hp + (size_t)tile0 * kBlockBytes, d.Mt, d.Vp / 128, d.Ip / kBN,
                                                    kbM, splits, ep, s_dw2, "tc_joiner_dW2_gemm", live))
        return rc;
    }
    cudaStream_t s_dw1 = fj.side(1);  // forked here: after dhidden
    // dW1[i, v] += sum_m dhidden[m, i] act(.)[m, v]: accumulator rows v, cols i; A = the J kept by the forward
    // pass (bulk copies) or, when it was too large to keep, rebuilt on the fly
    {
      StoreTransposedAtomicEpi ep{dW1, p.V, p.V, p.I};
      if (w.Jp) {
        BulkA a{w.Jp + (size_t)tile0 * kBlockBytes, d.Mt};
        if (int rc = launch_gemm_stream<kBN, kNStages, true, 0>(a, w.DHp, ct, d.Vp / 128, d.Ip / kBN, kbM, splits, ep, s_dw1,
                                                                "tc_joiner_dW1_gemm", live))
          return rc;
      } else {
        JointMnProducer a{p.am, p.lm, w.am_row, w.lm_row, row0, M, p.V, p.act};
        if (int rc = launch_gemm_stream<kBN, kNStages, true, 0>(a, w.DHp, ct, d.Vp / 128, d.Ip / kBN, kbM, splits, ep, s_dw1,
                                                                "tc_joiner_dW1_gemm"))
          return rc;
      }
    }
    // dh = dhidden W1: rows m, N = Vp, K = Ip  ->  bf16 rows; then the two segmented reductions
    {
      BulkA a{w.DHp, ct};
      StoreRowsBf16Epi ep{w.dh, d.Vp};
      if (int rc = launch_gemm_bstationary<kBN, kNStagesRes, kResSteps>(a, w.W1Tp, d.Vp / 128, ct, d.n_tiles_v, d.kbI, ep, stream,
                                                                       "tc_joiner_dh_gemm", live))
        return rc;
      const int64_t rows_live = (M - row0 < rows_pad) ? (M - row0) : rows_pad;
      {
        ProfScope prof("djoint_reduce_kernel", stream);
        const bool vec = (p.V % 4 == 0) && (((uintptr_t)p.am | (uintptr_t)p.lm | (uintptr_t)d_am | (uintptr_t)d_lm) % 16 == 0);
        // a frame whose band slots straddle two row chunks is finished by the second launch
        const bool am_acc = d.chunk < (int64_t)d.Mt * 128;
        const int rb = p.R <= 4 ? 4 : (p.R == 5 ? 5 : (p.R == 6 ? 6 : 8));
        if (!vec) launch_djoint_reduce<false, 4>(p, w.dh, d.Vp, row0, rows_live, M, am_acc, d_am, d_lm, stream);
        else if (rb == 4) launch_djoint_reduce<true, 4>(p, w.dh, d.Vp, row0, rows_live, M, am_acc, d_am, d_lm, stream);
        else if (rb == 5) launch_djoint_reduce<true, 5>(p, w.dh, d.Vp, row0, rows_live, M, am_acc, d_am, d_lm, stream);
        else if (rb == 6) launch_djoint_reduce<true, 6>(p, w.dh, d.Vp, row0, rows_live, M, am_acc, d_am, d_lm, stream);
        else launch_djoint_reduce<true, 8>(p, w.dh, d.Vp, row0, rows_live, M, am_acc, d_am, d_lm, stream);
      }
      if (int rc = check_launch("djoint reduce kernels")) return rc;
    }
    fj.join();  // the next chunk overwrites G and dhidden
  }
  return 0;
}

}  // namespace s2t
