// Tensor-core version of the simple-loss contractions (same maths as simple_loss.cu; replaces the
// bmm inside k2.rnnt_loss_smoothed, /root/reference/model/joiner/joiner.py:100-110).
//
//   forward   acc[b, t, s] = sum_c exp(am[b,t,c] - am_max) * exp(lm[b,s,c] - lm_max)
//             3xTF32 (fp32-level accuracy: these normalisers decide the prune ranges by an argmax);
//             A = exp(am - max) built on the fly (big | small), B = packed exp(lm - max) (big, small);
//             the epilogue thread owns one frame t and writes nrm / px / py columns s with the
//             k2 (B,S,T) layout, i.e. coalesced along t across the warp.
//   backward  d_am[b,t,c] = -exp(am - max) * sum_s W[b,s,t] exp(lm[b,s,c] - max)
//             d_lm[b,s,c] = -exp(lm - max) * sum_t W[b,s,t] exp(am[b,t,c] - max)
//             bf16 operands, MN-major (the contraction index is the row index of every packed operand),
//             batched over b; the one-hot terms are added by the scatter kernels of simple_loss.cu.
#include <stdlib.h>
#include "tc_gemm.cuh"

namespace s2t {

int simple_scatter_onehot(const float* occ_px, const float* occ_py, const int64_t* sym, const float* coef, int B,
                          int S, int T, int V, int blank, float* d_am, float* d_lm, cudaStream_t stream);

namespace {

using namespace tc;

// K-major contractions run as 2-CTA clusters: the B operand of a stage is loaded half by each CTA and multicast
constexpr int kPair = 1;  // CTA pairs (cta_group::2) measured a shade slower for this short batched contraction

// one warp per row: max over V.  For the lm rows (info != nullptr) lane 0 also records what the
// normaliser epilogue needs per symbol position: {max, lm[blank], lm[sym[b, s]]} as one 16-byte record.
__global__ void tc_row_max_kernel(const float* __restrict__ x, int64_t rows, int V, float* __restrict__ out,
                                  const int64_t* __restrict__ sym, int S, int blank, float4* __restrict__ info) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x % 32;
  const float* p = x + row * V;
  float m = kNegInf;
  for (int c = lane; c < V; c += 32) m = fmaxf(m, __ldg(p + c));
  m = warp_max(m);
  if (lane == 0) {
    out[row] = m;
    if (info) {
      const int64_t b = row / (S + 1);
      const int s = (int)(row % (S + 1));
      const float ls = (s < S) ? __ldg(p + (int)sym[b * S + s]) : 0.f;
      info[row] = make_float4(m, __ldg(p + blank), ls, 0.f);
    }
  }
}

// With the row maxima already known (by-products of the projection epilogue): the per-symbol-position record the
// normaliser epilogue needs, {max, lm[blank], lm[sym[b, s]]}, is three loads per row instead of a pass over the row.
__global__ void tc_lm_info_kernel(const float* __restrict__ lm, const float* __restrict__ lm_max, int64_t rows, int V,
                                  const int64_t* __restrict__ sym, int S, int blank, float4* __restrict__ info) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float* p = lm + row * V;
  const int64_t b = row / (S + 1);
  const int s = (int)(row % (S + 1));
  const float ls = (s < S) ? __ldg(p + (int)sym[b * S + s]) : 0.f;
  info[row] = make_float4(lm_max[row], __ldg(p + blank), ls, 0.f);
}

// K-major 3xTF32 A: rows = frames t of batch b, 32 vocabulary entries per k-step, value exp(am - am_max).
// Same thread mapping as RowCopyProducerF32: eight lanes read the 128 contiguous bytes of one row.
struct ExpRowProducerF32 {
  static constexpr bool kBulk = false;
  const float* x;   // (B, rows, V)
  const float* mx;  // (B, rows)
  int rows, V;
  // by-product for the backward contractions: exp(x - max) as a bf16 packed operand, batch b / row tile m at
  // row block b * tiles_per_batch + m, 64 columns per block (written while column tile 0 streams by)
  uint8_t* bf16_pack;
  int tiles_per_batch, pack_row_blocks;
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int c = lane & 7, rbase = warp * 4 + (lane >> 3);
    const bool vec = ((V & 3) == 0);
    const bool emit = bf16_pack != nullptr && pc.n_tile == 0 && pc.valid;
    const float* rowp[4];
    float sub[4];
    bool live[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = pc.m_tile * 128 + rbase + 32 * i;
      live[i] = r < rows;
      const int64_t g = (int64_t)pc.batch * rows + (live[i] ? r : 0);
      rowp[i] = x + g * V;
      sub[i] = live[i] ? __ldg(mx + g) : 0.f;
    }
    // three register buffers: the loads of k-steps it+1 and it+2 are in flight while step it is converted
    float4 buf[3][4];
    auto load = [&](float4 (&dst)[4], int ks) {
      const int k = ks * 32 + c * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (live[i] && vec && k + 4 <= V) {
          dst[i] = __ldg(reinterpret_cast<const float4*>(rowp[i] + k));
        } else {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (live[i] && k + j < V) ? __ldg(rowp[i] + k + j) : kNegInf;
          dst[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    const int off = ((c ^ (rbase & 7)) & 7) << 4;
    auto emit_stage = [&](const float4 (&cur)[4], int it) {
      pc.wait_empty(it);
      uint8_t* dst = pc.stage(it) + rbase * 128 + off;
      const int k = (pc.ks0 + it) * 32 + c * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
        float4 big, small;
        float* pb = &big.x;
        float* ps = &small.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = (live[i] && k + j < V) ? __expf(e[j] - sub[i]) : 0.f;
          pb[j] = round_tf32(p);
          ps[j] = round_tf32(p - pb[j]);
        }
        *reinterpret_cast<float4*>(dst + i * 32 * 128) = big;
        *reinterpret_cast<float4*>(dst + kBlockBytes + i * 32 * 128) = small;
        if (emit) {
          const int ks = pc.ks0 + it;
          uint8_t* blk = bf16_pack +
                         packed_block_index(pc.batch * tiles_per_batch + pc.m_tile, ks >> 1, pack_row_blocks) * kBlockBytes;
          *reinterpret_cast<uint2*>(blk + block_chunk_offset(rbase + 32 * i, (ks & 1) * 4 + (c >> 1)) + (c & 1) * 8) =
              make_uint2(pack_bf16x2(pb[0] + ps[0], pb[1] + ps[1]), pack_bf16x2(pb[2] + ps[2], pb[3] + ps[3]));
        }
      }
      pc.arrive_full(it);
    };
    load(buf[0], pc.ks0);
    if (pc.n_it > 1) load(buf[1], pc.ks0 + 1);
    for (int it = 0; it < pc.n_it; it += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (it + u < pc.n_it) {
          if (it + u + 2 < pc.n_it) load(buf[(u + 2) % 3], pc.ks0 + it + u + 2);
          emit_stage(buf[u], it + u);
        }
      }
    }
  }
};

// The same A operand for the 3xF16 contraction: 64 vocabulary entries per k-step, hi | lo f16 halves of
// scale * exp(am - am_max) (thread mapping of tc::RowSplitProducerF16), bf16 by-product unscaled.
struct ExpRowSplitProducerF16 {
  static constexpr bool kBulk = false;
  const float* x;   // (B, rows, V)
  const float* mx;  // (B, rows)
  int rows, V;
  float scale;
  uint8_t* bf16_pack;
  int tiles_per_batch, pack_row_blocks;
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int c = lane & 15, rbase = warp * 2 + (lane >> 4);
    const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool emit = bf16_pack != nullptr && pc.n_tile == 0 && pc.valid;
    const int off = rbase * 128 + ((((c >> 1) ^ (rbase & 7)) & 7) << 4) + (c & 1) * 8;
    float sub[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = pc.m_tile * 128 + rbase + 16 * i;
      sub[i] = r < rows ? __ldg(mx + (int64_t)pc.batch * rows + r) : 0.f;
    }
    auto load = [&](float4 (&dst)[4], int ks, int half) {
      const int k = ks * 64 + c * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = pc.m_tile * 128 + rbase + 16 * (half * 4 + j);
        const float* row = x + ((int64_t)pc.batch * rows + r) * V;
        if (r < rows && vec && k + 4 <= V) {
          dst[j] = __ldg(reinterpret_cast<const float4*>(row + k));
        } else {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (r < rows && k + e < V) ? __ldg(row + k + e) : kNegInf;
          dst[j] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    auto emit_half = [&](const float4 (&src)[4], int it, int half) {
      uint8_t* dst = pc.stage(it) + off;
      const int ks = pc.ks0 + it;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float s = sub[half * 4 + j];  // padding holds -inf -> exp = 0
        const float p[4] = {__expf(src[j].x - s), __expf(src[j].y - s), __expf(src[j].z - s), __expf(src[j].w - s)};
        const float v[4] = {p[0] * scale, p[1] * scale, p[2] * scale, p[3] * scale};
        uint2 hi, lo;
        split_f16x4(v, hi, lo);
        const int ro = (half * 4 + j) * 16 * 128;
        *reinterpret_cast<uint2*>(dst + ro) = hi;
        *reinterpret_cast<uint2*>(dst + kBlockBytes + ro) = lo;
        if (emit) {
          uint8_t* blk = bf16_pack + packed_block_index(pc.batch * tiles_per_batch + pc.m_tile, ks, pack_row_blocks) * kBlockBytes;
          *reinterpret_cast<uint2*>(blk + off + ro) = make_uint2(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]));
        }
      }
    };
    float4 q0[4], q1[4];
    load(q0, pc.ks0, 0);
    for (int it = 0; it < pc.n_it; ++it) {
      load(q1, pc.ks0 + it, 1);
      pc.wait_empty(it);
      emit_half(q0, it, 0);
      if (it + 1 < pc.n_it) load(q0, pc.ks0 + it + 1, 0);
      emit_half(q1, it, 1);
      pc.arrive_full(it);
    }
  }
};

// accumulator rows = frames t (ctx.m inside batch ctx.batch), columns = symbol positions s.
// Writes nrm and py (lane = frame, so every store instruction is one contiguous 128-byte row segment of the
// k2 (B, S+1, T) layout); px needs the gather am[b, t, sym[b, s]] and is finished by simple_px_kernel with
// the whole GPU instead of the four epilogue warps (measured: the gather inside this epilogue -- eight loads in flight
// per thread, four warps -- took the kernel from 0.057 to 0.105 ms; simple_px_kernel does it in 0.023).
struct SimpleEmitTcEpi {
  static constexpr int kScratchBytes = 0;
  const float* am;
  const float* am_max;
  const float4* lm_info;  // (B, S+1): {lm_max, lm[blank], lm[sym], -}
  int T, S, V, blank;
  float* py;   // (B, S+1, T)
  float* nrm;  // (B, S+1, T)
  float acc_scale;  // undoes the power-of-two pre-scale of the two operands
  struct State {
    float amx, am_blank;
    bool live;
  };
  __device__ void begin(State& st, const EpiCtx& ctx) const {
    const int b = ctx.batch, t = ctx.m;
    st.live = t < T;
    const int64_t r = (int64_t)b * T + (st.live ? t : 0);
    st.amx = st.live ? am_max[r] : 0.f;
    st.am_blank = st.live ? __ldg(am + r * V + blank) : 0.f;
  }
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State& st, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    if (!st.live || n > S) return;
    const int b = ctx.batch, t = ctx.m;
    const float4* info = lm_info + (int64_t)b * (S + 1);
    float* nrow = nrm + ((int64_t)b * (S + 1) + n) * T + t;
    float* prow = py + ((int64_t)b * (S + 1) + n) * T + t;
    // the per-column lm terms are fetched sixteen at a time ahead of the stores they feed: interleaved with the
    // stores every load would expose its own L2 round trip (the producers' streaming loads own the small L1)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float lx[16], ly[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = n + g * 16 + j;
        const float4 li = __ldg(info + (c <= S ? c : S));
        lx[j] = li.x;
        ly[j] = li.y;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int jj = g * 16 + j;
        if (n + jj <= S) {
          const float nv = logf(acc[jj] * acc_scale + FLT_MIN) + lx[j] + st.amx;
          nrow[(int64_t)jj * T] = nv;
          prow[(int64_t)jj * T] = st.am_blank + ly[j] - nv;
        }
      }
    }
  }
};

// px[b, s, t] = am[b, t, sym[b, s]] + lm[b, s, sym[b, s]] - nrm[b, s, t]   (-inf at t = T_b and in column T)
// Tiled version: one CTA per (utterance, 32 frames).  Warp w gathers the S symbol entries of am row t0 + w (one 2 KB
// row: a handful of lines per warp load, instead of 32 rows per load when consecutive lanes are consecutive frames)
// into a shared-memory tile; the tile is then read frame-fastest, so that nrm is read and px written in 128-byte
// segments of the k2 layout (t contiguous).  Shared memory: 32 x (S | 1) floats.
__global__ void __launch_bounds__(1024) simple_px_tiled_kernel(const float* __restrict__ am, const float4* __restrict__ lm_info,
                                                               const float* __restrict__ nrm, const int64_t* __restrict__ sym,
                                                               const int64_t* __restrict__ boundary, int B, int T, int S,
                                                               int V, float* __restrict__ px) {
  extern __shared__ float tile[];  // [32][ld]
  const int ld = S | 1;
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tb = boundary ? (int)boundary[4 * b + 3] : T;
  const int64_t* sy = sym + (int64_t)b * S;
  {
    const int t = t0 + warp;
    if (t < T) {
      const float* row = am + ((int64_t)b * T + t) * V;
      for (int s = lane; s < S; s += 32) {
        const int c = min(max((int)sy[s], 0), V - 1);
        tile[warp * ld + s] = __ldg(row + c);
      }
    }
  }
  __syncthreads();
  // warp w now owns the symbol positions w, w + 32, ...; lane = frame inside the tile
  const int t = t0 + lane;
  for (int s = warp; s < S; s += 32) {
    const int64_t row = (int64_t)b * (S + 1) + s;
    float v = kNegInf;
    if (t < T && t != Tb) v = tile[lane * ld + s] + __ldg(lm_info + row).z - __ldg(nrm + row * T + t);
    if (t < T) px[((int64_t)b * S + s) * (T + 1) + t] = v;
    if (t0 + 32 >= T && lane == 0) px[((int64_t)b * S + s) * (T + 1) + T] = kNegInf;  // column T
  }
}

__global__ void simple_px_kernel(const float* __restrict__ am, const float4* __restrict__ lm_info,
                                 const float* __restrict__ nrm, const int64_t* __restrict__ sym,
                                 const int64_t* __restrict__ boundary, int B, int T, int S, int V,
                                 float* __restrict__ px) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * S * (T + 1);
  if (i >= total) return;
  const int t = (int)(i % (T + 1));
  const int64_t bs = i / (T + 1);
  const int s = (int)(bs % S), b = (int)(bs / S);
  float v = kNegInf;
  const int Tb = boundary ? (int)boundary[4 * b + 3] : T;
  if (t < T && t != Tb) {
    const int c = (int)sym[bs];
    const int64_t row = (int64_t)b * (S + 1) + s;
    v = __ldg(am + ((int64_t)b * T + t) * V + c) + __ldg(lm_info + row).z - __ldg(nrm + row * T + t);
  }
  px[i] = v;
}

// W[b,s,t] = coef_b (occ_px + occ_py) exp(am_max + lm_max - nrm) -> bf16 packed twice:
//   Wst: rows (b, s) cols t        Wts: rows (b, t) cols s
// One CTA per (b, 64 x 64 tile): the tile is computed once (reads coalesced along t), parked in shared memory
// as bf16 and written out in both orientations as whole 16-byte chunks.
__global__ void __launch_bounds__(256) simple_w_packed_kernel(const float* __restrict__ occ_px,
                                                              const float* __restrict__ occ_py,
                                                              const float* __restrict__ nrm,
                                                              const float* __restrict__ am_max,
                                                              const float* __restrict__ lm_max,
                                                              const float* __restrict__ coef, int B, int S, int T,
                                                              int Spad, int Tpad, uint8_t* __restrict__ Wst,
                                                              uint8_t* __restrict__ Wts) {
  __shared__ __nv_bfloat16 w[64][66];  // odd word stride: column reads are conflict-free
  const int b = blockIdx.z, s0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // t lane, s lane
  const float cf = coef[b];
  const int t = t0 + tx;
  const float amx = t < T ? am_max[(int64_t)b * T + t] : 0.f;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int sl = ty + 4 * i, s = s0 + sl;
    float v = 0.f;
    if (s <= S && t < T) {
      const int64_t o = ((int64_t)b * (S + 1) + s) * T + t;
      float g = occ_py[o];
      if (s < S) g += occ_px[((int64_t)b * S + s) * (T + 1) + t];
      if (g != 0.f) v = cf * g * expf(amx + lm_max[(int64_t)b * (S + 1) + s] - nrm[o]);
    }
    w[sl][tx] = __float2bfloat16(v);
  }
  __syncthreads();
  // 64 rows x 8 chunks per orientation: two chunks per thread and orientation
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = threadIdx.x + 256 * i;  // 0..511
    const int row = q >> 3, ck = q & 7;
    {
      // Wst: row (b, s0 + row), columns t0 + 8 ck .. + 7
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = *reinterpret_cast<const uint32_t*>(&w[row][ck * 8 + 2 * j]);
      const int64_t r = (int64_t)b * Spad + s0 + row;
      uint8_t* blk = Wst + packed_block_index((int)(r >> 7), t0 >> 6, (int)(((int64_t)B * Spad) >> 7)) * kBlockBytes;
      *reinterpret_cast<uint4*>(blk + block_chunk_offset((int)(r & 127), ck)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    {
      // Wts: row (b, t0 + row), columns s0 + 8 ck .. + 7 (a column of the tile)
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint16_t lo = *reinterpret_cast<const uint16_t*>(&w[ck * 8 + 2 * j][row]);
        const uint16_t hi = *reinterpret_cast<const uint16_t*>(&w[ck * 8 + 2 * j + 1][row]);
        o[j] = (uint32_t)lo | ((uint32_t)hi << 16);
      }
      const int64_t r = (int64_t)b * Tpad + t0 + row;
      uint8_t* blk = Wts + packed_block_index((int)(r >> 7), s0 >> 6, (int)(((int64_t)B * Tpad) >> 7)) * kBlockBytes;
      *reinterpret_cast<uint4*>(blk + block_chunk_offset((int)(r & 127), ck)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// out[b, r, c] = -exp(x[b, r, c] - max[b, r]) * acc      accumulator rows r (inside batch), cols c.
// Transposed through shared memory so that x is read and out written in 64-byte row segments.
struct GradExpEpi {
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  const float* x;
  const float* mx;
  int rows, V;
  float* out;
  struct State { float sub[4]; };  // row maxima of the four rows this lane serves after the transpose
  __device__ void begin(State& st, const EpiCtx& ctx) const {
    const int lane = ctx.t & 31, m0 = ctx.m - lane;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int r = m0 + pass * 8 + (lane & 7);
      st.sub[pass] = r < rows ? mx[(int64_t)ctx.batch * rows + r] : 0.f;
    }
  }
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State& st, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    const int lane = ctx.t & 31, m0 = ctx.m - lane;
    const bool vec = ((V & 3) == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    // the eight x pieces this lane will need after the transpose, all requested before the first one is used
    float4 xs[8];
    if (vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i & 3) * 8 + (lane & 7), col = n + (i >> 2) * 16 + (lane >> 3) * 4;
        xs[i] = (m < rows && col < V) ? __ldg(reinterpret_cast<const float4*>(x + ((int64_t)ctx.batch * rows + m) * V + col))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    warp_transposed_chunk(ctx, acc, V - n, [&](int r, int c, float4 v) {
      const int m = m0 + r, col = n + c;
      if (m >= rows || col >= V) return;
      const float sub = st.sub[r >> 3];
      const int64_t o = ((int64_t)ctx.batch * rows + m) * V + col;
      if (vec) {
        const float4 xv = xs[(c >> 4) * 4 + (r >> 3)];
        *reinterpret_cast<float4*>(out + o) = make_float4(-__expf(xv.x - sub) * v.x, -__expf(xv.y - sub) * v.y,
                                                          -__expf(xv.z - sub) * v.z, -__expf(xv.w - sub) * v.w);
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < V) out[o + j] = -__expf(__ldg(x + o + j) - sub) * e[j];
      }
    });
  }
};

struct SimpleTcDims {
  int Spad, Tpad, Vp;  // S+1 and T padded to 128, V padded to 256
  int kb32, kb64;      // vocabulary blocks of 32 (fp32) / 64 (bf16, = Vp / 64)
  size_t lm_f32, am_bf16, lm_bf16, wst, wts, info;
};

SimpleTcDims simple_tc_dims(int B, int T, int S, int V) {
  SimpleTcDims d;
  d.Spad = ((S + 1 + 127) / 128) * 128;
  d.Tpad = ((T + 127) / 128) * 128;
  d.Vp = ((V + 255) / 256) * 256;
  d.kb32 = (V + 31) / 32;
  d.kb64 = d.Vp / 64;
  d.lm_f32 = (size_t)B * (d.Spad / 128) * d.kb32 * kBlockBytes;
  d.am_bf16 = (size_t)B * (d.Tpad / 128) * d.kb64 * kBlockBytes;
  d.lm_bf16 = (size_t)B * (d.Spad / 128) * d.kb64 * kBlockBytes;
  d.wst = (size_t)B * (d.Spad / 128) * (d.Tpad / 64) * kBlockBytes;
  d.wts = (size_t)B * (d.Tpad / 128) * (d.Spad / 64) * kBlockBytes;
  d.info = (((size_t)B * (S + 1) * sizeof(float4)) + 1023) / 1024 * 1024;
  return d;
}

}  // namespace

size_t simple_tc_workspace_bytes(int B, int T, int S, int V) {
  SimpleTcDims d = simple_tc_dims(B, T, S, V);
  return 2 * d.lm_f32 + d.am_bf16 + d.lm_bf16 + d.wst + d.wts + d.info + 4096;
}

// The lm side of the normaliser, separately: {max, lm[blank], lm[sym]} per symbol position and the hi / lo split of
// exp(lm - max) (the B operand).  It depends on the predictor-side projection only, so a caller that launches that
// projection on a second stream issues this right behind it -- under the (three times larger) encoder-side projection.
int simple_prep_lm_tc(const float* lm, const float* lm_max, const int64_t* sym, int B, int T, int S, int V, int blank,
                      void* ws, cudaStream_t stream) {
  SimpleTcDims d = simple_tc_dims(B, T, S, V);
  uint8_t* lm_big = (uint8_t*)ws;
  uint8_t* lm_small = lm_big + d.lm_f32;
  float4* lm_info = (float4*)((uint8_t*)ws + 2 * d.lm_f32 + d.am_bf16 + d.lm_bf16 + d.wst + d.wts);
  const int64_t rows_lm = (int64_t)B * (S + 1);
  {
    ProfScope prof("lm_info_kernel", stream);
    tc_lm_info_kernel<<<(unsigned)((rows_lm + 255) / 256), 256, 0, stream>>>(lm, lm_max, rows_lm, V, sym, S, blank, lm_info);
  }
  if (int rc = check_launch("lm_info_kernel")) return rc;
  uint8_t* am_p = lm_small + d.lm_f32;
  uint8_t* lm_p = am_p + d.am_bf16;
  const bool f16 = getenv("S2T_B200_SIMPLE_TF32") == nullptr;
  const int kstep = f16 ? 64 : 32;
  const int ksteps = (V + kstep - 1) / kstep;
  if (ksteps * kstep < d.Vp) cudaMemsetAsync(lm_p, 0, d.lm_bf16, stream);  // vocabulary padding the k-steps never visit
  if (f16) {
    constexpr float kScale = 4096.f;
    PackSpec ps{lm, (int64_t)(S + 1) * V, V, B, S + 1, d.Spad, V, ksteps, lm_max};
    return pack_f16_split(ps, kScale, lm_big, lm_small, lm_p, stream);
  }
  PackSpec ps{lm, (int64_t)(S + 1) * V, V, B, S + 1, d.Spad, V, d.kb32, lm_max};
  return pack_f32_split(ps, lm_big, lm_small, stream, lm_p, d.kb64);
}

// ready: 0 = nothing known; 1 = am_max / lm_max hold the row maxima; 2 = also simple_prep_lm_tc has run on this workspace
int simple_logprobs_tc(const float* am, const float* lm, const int64_t* sym, const int64_t* boundary, int B, int T,
                       int S, int V, int blank, float* am_max, float* lm_max, float* px, float* py, float* nrm,
                       void* ws, int ready, cudaStream_t stream) {
  const bool row_max_ready = ready != 0, lm_ready = ready == 2;
  SimpleTcDims d = simple_tc_dims(B, T, S, V);
  uint8_t* lm_big = (uint8_t*)ws;
  uint8_t* lm_small = lm_big + d.lm_f32;
  float4* lm_info = (float4*)((uint8_t*)ws + 2 * d.lm_f32 + d.am_bf16 + d.lm_bf16 + d.wst + d.wts);
  const int wpb = 8;
  const int64_t rows_am = (int64_t)B * T, rows_lm = (int64_t)B * (S + 1);
  // lm first: its hi/lo split (the B operand of the normaliser) then runs on a side stream next to the am row maxima
  if (lm_ready) {
  } else if (row_max_ready) {
    ProfScope prof("lm_info_kernel", stream);
    tc_lm_info_kernel<<<(unsigned)((rows_lm + 255) / 256), 256, 0, stream>>>(lm, lm_max, rows_lm, V, sym, S, blank, lm_info);
  } else {
    ProfScope prof("row_max_kernel", stream);
    tc_row_max_kernel<<<(unsigned)((rows_lm + wpb - 1) / wpb), wpb * 32, 0, stream>>>(lm, rows_lm, V, lm_max, sym, S,
                                                                                      blank, lm_info);
  }
  if (int rc = check_launch("row_max_kernel")) return rc;
  // exp(am - max) and exp(lm - max) as bf16 operands of the backward contractions are by-products of this pass
  uint8_t* am_p = lm_small + d.lm_f32;
  uint8_t* lm_p = am_p + d.am_bf16;
  const bool f16 = getenv("S2T_B200_SIMPLE_TF32") == nullptr;  // default: 3xF16 split; 3xTF32 stays selectable
  const int kstep = f16 ? 64 : 32;
  const int ksteps = (V + kstep - 1) / kstep;
  if (ksteps * kstep < d.Vp) {  // vocabulary padding the k-steps never visit
    cudaMemsetAsync(am_p, 0, d.am_bf16 + (lm_ready ? 0 : d.lm_bf16), stream);
  }
  MnDebug extra;
  extra.b_small = lm_small;
  extra.b_batch_off = d.Spad / 128;
  auto row_max_am = [&]() {
    if (row_max_ready) return;  // am_max came out of the projection epilogue
    ProfScope prof("row_max_kernel", stream);
    tc_row_max_kernel<<<(unsigned)((rows_am + wpb - 1) / wpb), wpb * 32, 0, stream>>>(am, rows_am, V, am_max, nullptr, S,
                                                                                      blank, nullptr);
  };
  if (f16) {
    // both operands are in (0, 1]: scaled by 2^12 so that entries down to ~3e-5 keep a normal lo half; the product
    // is scaled back (2^-24, exact) before the log
    constexpr float kScale = 4096.f;
    PackSpec ps{lm, (int64_t)(S + 1) * V, V, B, S + 1, d.Spad, V, ksteps, lm_max};
    if (!lm_ready) {
      ForkJoin fj(stream);
      if (int rc = pack_f16_split(ps, kScale, lm_big, lm_small, lm_p, fj.side(0))) return rc;
      row_max_am();
    }  // joined
    if (int rc = check_launch("row_max_kernel")) return rc;
    ExpRowSplitProducerF16 a{am, am_max, T, V, kScale, am_p, d.Tpad / 128, B * (d.Tpad / 128)};
    SimpleEmitTcEpi ep{am, am_max, lm_info, T, S, V, blank, py, nrm, 1.f / (kScale * kScale)};
    if (int rc = launch_gemm_stream<128, 3, false, 3, kPair>(a, lm_big, B * (d.Spad / 128), d.Tpad / 128, d.Spad / 128, ksteps,
                                                             1, ep, stream, "tc_simple_normaliser_gemm_3xf16", extra, B))
      return rc;
  } else {
    PackSpec ps{lm, (int64_t)(S + 1) * V, V, B, S + 1, d.Spad, V, d.kb32, lm_max};
    row_max_am();
    if (int rc = check_launch("row_max_kernel")) return rc;
    if (!lm_ready)
      if (int rc = pack_f32_split(ps, lm_big, lm_small, stream, lm_p, d.kb64)) return rc;
    ExpRowProducerF32 a{am, am_max, T, V, am_p, d.Tpad / 128, B * (d.Tpad / 128)};
    SimpleEmitTcEpi ep{am, am_max, lm_info, T, S, V, blank, py, nrm, 1.f};
    if (int rc = launch_gemm_stream<128, 3, false, 2, kPair>(a, lm_big, B * (d.Spad / 128), d.Tpad / 128, d.Spad / 128, d.kb32,
                                                             1, ep, stream, "tc_simple_normaliser_gemm_3xtf32", extra, B))
      return rc;
  }
  const int64_t total = (int64_t)B * S * (T + 1);
  if (total > 0) {
    ProfScope prof("simple_px_kernel", stream);
    const size_t smem = (size_t)32 * (S | 1) * sizeof(float);
    if (smem <= 48 * 1024 && B <= 65535 && getenv("S2T_B200_PX_FLAT") == nullptr) {
      simple_px_tiled_kernel<<<dim3((unsigned)((T + 31) / 32), (unsigned)B), 1024, smem, stream>>>(am, lm_info, nrm, sym, boundary,
                                                                                                  B, T, S, V, px);
    } else {
      simple_px_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(am, lm_info, nrm, sym, boundary, B, T, S, V, px);
    }
  }
  return check_launch("simple_px_kernel");
}

int simple_backward_tc(const float* am, const float* lm, const int64_t* sym, const float* am_max,
                       const float* lm_max, const float* nrm, const float* occ_px, const float* occ_py,
                       const float* coef, int B, int T, int S, int V, int blank, void* ws, float* d_am, float* d_lm,
                       cudaStream_t stream) {
  SimpleTcDims d = simple_tc_dims(B, T, S, V);
  uint8_t* p = (uint8_t*)ws + 2 * d.lm_f32;
  uint8_t* am_p = p; p += d.am_bf16;
  uint8_t* lm_p = p; p += d.lm_bf16;
  uint8_t* Wst = p; p += d.wst;
  uint8_t* Wts = p;
  {
    ProfScope prof("simple_w_kernel", stream);
    const dim3 grid((unsigned)(d.Tpad / 64), (unsigned)(d.Spad / 64), (unsigned)B);
    simple_w_packed_kernel<<<grid, 256, 0, stream>>>(occ_px, occ_py, nrm, am_max, lm_max, coef, B, S, T, d.Spad, d.Tpad,
                                                     Wst, Wts);
  }
  if (int rc = check_launch("simple_w_packed_kernel")) return rc;
  // am_p / lm_p = bf16 exp(am - max), exp(lm - max): written by simple_logprobs_tc into the SAME workspace
  ForkJoin fj(stream);  // the two gradient contractions are independent: d_lm runs on a side stream
  cudaStream_t s_lm = fj.side(0);
  // d_am: rows t, cols c, contraction over s (rows of Wst and of lm_p)
  {
    BulkA a{Wst, B * (d.Spad / 128)};
    GradExpEpi ep{am, am_max, T, V, d_am};
    MnDebug extra;
    extra.a_batch_off = d.Spad / 64;
    extra.b_batch_off = d.Spad / 64;
    if (int rc = launch_gemm_stream<256, 4, true, 0>(a, lm_p, B * (d.Spad / 128), d.Tpad / 128, d.Vp / 256, d.Spad / 64, 1,
                                                     ep, stream, "tc_simple_d_am_gemm", extra, B))
      return rc;
  }
  // d_lm: rows s, cols c, contraction over t (rows of Wts and of am_p)
  {
    BulkA a{Wts, B * (d.Tpad / 128)};
    GradExpEpi ep{lm, lm_max, S + 1, V, d_lm};
    MnDebug extra;
    extra.a_batch_off = d.Tpad / 64;
    extra.b_batch_off = d.Tpad / 64;
    if (int rc = launch_gemm_stream<256, 4, true, 0>(a, am_p, B * (d.Tpad / 128), d.Spad / 128, d.Vp / 256, d.Tpad / 64, 1,
                                                     ep, s_lm, "tc_simple_d_lm_gemm", extra, B))
      return rc;
  }
  fj.join();
  return simple_scatter_onehot(occ_px, occ_py, sym, coef, B, S, T, V, blank, d_am, d_lm, stream);
}

}  // namespace s2t
