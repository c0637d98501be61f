// bf16 x bf16 -> fp32 contractions on the 5th-generation tensor cores.
//
//     K-major  (kMn = false):  C[m, n] (+)= sum_k A[m, k]  * B[n, k]     operands stored (rows, K)
//     MN-major (kMn = true):   C[m, n] (+)= sum_k At[k, m] * Bt[k, n]    operands stored (K, rows)
//
// One CTA owns a 128 x BN output tile whose accumulator lives in TMEM; a single elected thread
// issues tcgen05.mma (M=128, N=BN, K=16) over 64-deep K steps that arrive in shared memory
// through a kStages-deep mbarrier ring.  Warp roles (192 threads):
//   warp 0      bulk-copy issuer (cp.async.bulk, TMA engine) for the packed operands
//   warp 1      TMEM allocation + MMA issue + commits
//   warps 2..5  (a) optional on-the-fly A producer -- the threads build the A stage (e.g.
//               act(am + lm[ranges]) -> bf16) directly in the swizzled smem image, so the operand
//               never exists in HBM;  (b) epilogue: tcgen05.ld the accumulator (one TMEM lane = one
//               output row per thread) and hand 32-column chunks to the epilogue functor.
//
// Both modes read the SAME packed format (tc_prims.cuh): 128 x 64 blocks, 128 B per row, 16-byte
// chunks XOR-swizzled by (row & 7).  For a K-major operand the block rows are the operand's
// M/N index; for an MN-major operand the block rows are the contraction index and the 64
// columns one "group" of the M/N index, so an activation written once as (rows = m, cols = i)
// serves as K-major A of a row-wise contraction AND as MN-major operand of a reduction over m
// (the weight-gradient contractions) without a transposed copy.
//
// smem stage layout
//   K-major :  A block [128 rows x 128 B]            | B blocks [BN rows x 128 B]
//   MN-major:  A groups 2 x [64 k-rows x 128 B]      | B groups (BN/64) x [64 k-rows x 128 B]
//
// Functor contracts
//   struct ASrc { static constexpr bool kBulk;
//       // kBulk : const uint8_t* packed; int row_blocks;   (row blocks of the packed array)
//       // !kBulk: template <class W, class A> __device__ void run(uint8_t* smem, int stage_bytes, int stages,
//       //             int m_tile, int ks0, int n_it, int t, int batch, W wait_empty, A arrive_full) const;
//       //         must, for it in [0, n_it): wait_empty(it); fill stage it % stages; arrive_full(it).
//   };
//   struct Epi {
//     struct State {...};                       // per-thread (= per output row) running state
//     __device__ void begin(State&, const EpiCtx&) const;
//     __device__ void chunk(State&, const EpiCtx&, int n, const float (&acc)[32]) const;   // 32 columns from n
//     __device__ void end(State&, const EpiCtx&) const;
//   };
// During the epilogue all MMAs of the CTA have completed, so the operand ring is free:
// EpiCtx::scratch points at it (kStages * stage bytes) for CTA-level reductions; epi_sync()
// is a barrier over the 128 epilogue threads.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace s2t {
namespace tc {

struct EpiCtx {
  int m;        // global output row of this thread
  int t;        // epilogue thread index, 0..127
  int row;      // row inside the 128-row tile
  int m_tile, n_tile, split, batch;
  uint8_t* scratch;
  int scratch_bytes;
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct BulkA {
  static constexpr bool kBulk = true;
  const uint8_t* packed;
  int row_blocks;
};

constexpr int kGemmThreads = 192;
constexpr int kGroupBytes = 64 * 128;  // one MN-major group: 64 k-rows x 64 elements

// kKind: 0 = bf16 operands; 1 = tf32 (fp32 in smem, single pass); 2 = 3xTF32: every fp32 operand is held
// as big = rn_tf32(x) and small = x - big, and D += A_big B_big + A_big B_small + A_small B_big, which
// recovers fp32-level accuracy (error ~2^-21) on the tensor cores.
template <int BN, int kStages, int kKind = 0>
constexpr size_t gemm_stream_smem_bytes() {
  return (size_t)kStages * (kKind == 2 ? 2 : 1) * (kBlockBytes + (BN / 128) * kBlockBytes) + 1024 /*align*/ +
         256 /*barriers*/;
}

// MN-major smem descriptor: 8-row (K) groups 1024 B apart, 64-element (MN) groups lbo_bytes apart
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

struct MnDebug {  // descriptor knobs (kept as kernel arguments so a test can probe the encoding)
  uint32_t lbo_bytes = kGroupBytes;
  uint32_t sbo_bytes = 1024;
  uint32_t k_advance_bytes = 2048;  // 16 k-rows
  const uint8_t* b_small = nullptr;  // kKind == 2: packed residuals of B
  // batched launches (blockIdx.z = batch * k_splits + split): operand offsets per batch, in 128-row
  // blocks for K-major operands and in 64-row k-steps for MN-major operands
  int a_batch_off = 0;
  int b_batch_off = 0;
};

template <int BN, int kStages, bool kMn, int kKind, class ASrc, class Epi>
__global__ void __launch_bounds__(kGemmThreads, kKind == 2 ? 1 : ((BN == 128 && ASrc::kBulk) ? 3 : 2))
gemm_stream_kernel(ASrc asrc, const uint8_t* __restrict__ b_packed, int b_row_blocks, int k_steps, int k_splits,
                   Epi epi, MnDebug mn) {
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
  constexpr int kParts = kKind == 2 ? 2 : 1;
  constexpr int kABytes = kParts * kBlockBytes;
  constexpr int kBPart = (BN / 128) * kBlockBytes;
  constexpr int kBBytes = kParts * kBPart;
  constexpr int kStageBytes = kABytes + kBBytes;
  static_assert(kKind != 2 || !ASrc::kBulk, "3xTF32 expects an on-the-fly A producer that writes big|small");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int batch = blockIdx.z / k_splits, split = blockIdx.z % k_splits;
  const int per = (k_steps + k_splits - 1) / k_splits;
  const int ks0 = split * per;
  const int ks1 = min(k_steps, ks0 + per);
  const int n_it = ks1 - ks0;
  if (n_it <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], ASrc::kBulk ? 1 : 1 + 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1, ks = ks0 + it;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full[s], (ASrc::kBulk ? kABytes : 0) + kBBytes);
        if constexpr (!kMn) {
          if constexpr (ASrc::kBulk) {
            bulk_copy_g2s(sa, asrc.packed + packed_block_index(m_tile + batch * mn.a_batch_off, ks, asrc.row_blocks) *
                                                kBlockBytes,
                          kABytes, &full[s]);
          }
          const size_t boff = packed_block_index(n_tile * (BN / 128) + batch * mn.b_batch_off, ks, b_row_blocks) *
                              kBlockBytes;
          bulk_copy_g2s(sb, b_packed + boff, kBPart, &full[s]);
          if constexpr (kKind == 2) bulk_copy_g2s(sb + kBPart, mn.b_small + boff, kBPart, &full[s]);
        } else {
          // k-step ks = 64 contraction rows = half of row block ks >> 1; group g = 64 columns = column block
          if constexpr (ASrc::kBulk) {
            const int ka = ks + batch * mn.a_batch_off;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              bulk_copy_g2s(sa + g * kGroupBytes,
                            asrc.packed + packed_block_index(ka >> 1, m_tile * 2 + g, asrc.row_blocks) * kBlockBytes +
                                (size_t)(ka & 1) * kGroupBytes,
                            kGroupBytes, &full[s]);
            }
          }
          const int kb2 = ks + batch * mn.b_batch_off;
#pragma unroll
          for (int g = 0; g < BN / 64; ++g) {
            bulk_copy_g2s(sb + g * kGroupBytes,
                          b_packed + packed_block_index(kb2 >> 1, n_tile * (BN / 64) + g, b_row_blocks) * kBlockBytes +
                              (size_t)(kb2 & 1) * kGroupBytes,
                          kGroupBytes, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // kKind 0: bf16 operands (64 per 128-byte row);  1: tf32 (fp32 in smem, 32 per row, K-major only)
      static_assert(kKind == 0 || !kMn, "tf32 operands are K-major only");
      uint32_t idesc = kKind == 0 ? umma_idesc_bf16(128, BN) : umma_idesc_tf32(128, BN);
      if (kMn) idesc |= (1u << 15) | (1u << 16);  // A and B are MN-major
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
          uint64_t da, db;
          if constexpr (!kMn) {
            da = umma_smem_desc(sa + k4 * kUmmaK * 2);
            db = umma_smem_desc(sb + k4 * kUmmaK * 2);
          } else {
            da = umma_smem_desc_mn(sa + k4 * mn.k_advance_bytes, mn.lbo_bytes, mn.sbo_bytes);
            db = umma_smem_desc_mn(sb + k4 * mn.k_advance_bytes, mn.lbo_bytes, mn.sbo_bytes);
          }
          if constexpr (kKind == 0) {
            umma_bf16(tmem_base, da, db, idesc, (it > 0) || (k4 > 0));
          } else if constexpr (kKind == 1) {
            umma_tf32(tmem_base, da, db, idesc, (it > 0) || (k4 > 0));
          } else {
            const uint64_t da_s = umma_smem_desc(sa + kBlockBytes + k4 * kUmmaK * 2);
            const uint64_t db_s = umma_smem_desc(sb + kBPart + k4 * kUmmaK * 2);
            umma_tf32(tmem_base, da_s, db, idesc, (it > 0) || (k4 > 0));  // small terms first
            umma_tf32(tmem_base, da, db_s, idesc, true);
            umma_tf32(tmem_base, da, db, idesc, true);
          }
        }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    const int t = (warp - 2) * 32 + lane;
    if constexpr (!ASrc::kBulk) {
      asrc.run(
          smem, kStageBytes, kStages, m_tile, ks0, n_it, t, batch,
          [&](int it) { mbar_wait(&empty[it % kStages], ((it / kStages) & 1) ^ 1); },
          [&](int it) {
            fence_proxy_async_smem();
            mbar_arrive(&full[it % kStages]);
          });
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    EpiCtx ctx;
    ctx.row = quarter * 32 + lane;
    ctx.m = m_tile * 128 + ctx.row;
    ctx.t = t;
    ctx.m_tile = m_tile;
    ctx.n_tile = n_tile;
    ctx.split = split;
    ctx.batch = batch;
    ctx.scratch = smem;
    ctx.scratch_bytes = kStages * kStageBytes;
    typename Epi::State st;
    epi.begin(st, ctx);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, v);
      epi.chunk(st, ctx, n_tile * BN + c * 32, v);
    }
    epi.end(st, ctx);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

template <int BN, int kStages, bool kMn, int kKind, class ASrc, class Epi>
int launch_gemm_stream(const ASrc& asrc, const uint8_t* b_packed, int b_row_blocks, int m_tiles, int n_tiles,
                       int k_steps, int k_splits, const Epi& epi, cudaStream_t stream, const char* what,
                       MnDebug mn = MnDebug(), int batches = 1) {
  if (m_tiles <= 0 || n_tiles <= 0 || k_steps <= 0) return 0;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > k_steps) k_splits = k_steps;
  auto kern = gemm_stream_kernel<BN, kStages, kMn, kKind, ASrc, Epi>;
  constexpr size_t smem = gemm_stream_smem_bytes<BN, kStages, kKind>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString(e));
      return 2;
    }
    configured = true;
  }
  dim3 grid(n_tiles, m_tiles, k_splits * batches);
  S2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", what);
  ProfScope prof(what, stream);
  kern<<<grid, kGemmThreads, smem, stream>>>(asrc, b_packed, b_row_blocks, k_steps, k_splits, epi, mn);
  return check_launch(what);
}

// ---- epilogues --------------------------------------------------------------------------------
// C[m * ldc + n] = acc (or += with atomics when partial sums from several splits / CTAs meet)
struct StoreRowMajorEpi {
  float* C;
  int64_t ldc;
  int M, N;
  bool atomic;
  const float* bias = nullptr;  // added per column when not atomic
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    const int m = ctx.m;
    if (m >= M) return;
    float* row = C + (int64_t)m * ldc;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (n + j < N) {
        if (atomic) atomicAdd(row + n + j, acc[j]);
        else row[n + j] = acc[j] + (bias ? __ldg(bias + n + j) : 0.f);
      }
    }
  }
};

// ---- packing ----------------------------------------------------------------------------------
// fp32 src(r, k) = src[r * row_stride + k * col_stride]  ->  bf16 packed operand of row_blocks x k_blocks
// blocks, zero padded beyond (rows, K).
int pack_operand(const float* src, int64_t row_stride, int64_t col_stride, int rows, int K, int row_blocks,
                 int k_blocks, uint8_t* dst, cudaStream_t stream);
// same block geometry with fp32 elements (32 per 128-byte row) for the tf32 contractions; k_blocks of 32
// part 0: rn_tf32(x);  part 1: residual x - rn_tf32(x) (itself rounded to tf32)
int pack_operand_f32(const float* src, int64_t row_stride, int rows, int K, int row_blocks, int k_blocks,
                     int part, uint8_t* dst, cudaStream_t stream);

// Batched packing with an optional fused exp: element (batch, r, k) = f(src[batch*batch_stride + r*row_stride + k])
// with f(x) = exp(x - row_sub[batch*rows + r]) when row_sub != nullptr.  Every batch is padded to rows_pad
// (a multiple of 128) rows, so the batches form one tall packed operand and never share a block.
struct PackSpec {
  const float* src;
  int64_t batch_stride, row_stride;
  int batches, rows, rows_pad, K;
  int k_blocks;          // column blocks of the packed operand (64 bf16 / 32 fp32 elements each)
  const float* row_sub;  // optional per-row value subtracted before exp
};
int pack_bf16(const PackSpec& p, uint8_t* dst, cudaStream_t stream);
int pack_f32_split(const PackSpec& p, uint8_t* dst_big, uint8_t* dst_small, cudaStream_t stream);

// On-the-fly K-major A for tf32: copies 128 rows x 32 fp32 of a row-major matrix into the swizzled stage.
struct RowCopyProducerF32 {
  static constexpr bool kBulk = false;
  const float* x;
  int64_t ld;
  int64_t M;
  int K;
  bool split;  // also write the residual block right after the big block (3xTF32)
  template <class W, class A>
  __device__ void run(uint8_t* smem, int stage_bytes, int stages, int m_tile, int ks0, int n_it, int t, int,
                      W wait_empty, A arrive_full) const {
    const int64_t m = (int64_t)m_tile * 128 + t;
    const bool live = m < M;
    const float* row = x + (live ? m * ld : 0);
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    float4 cur[8], nxt[8];
    auto load = [&](float4 (&dst)[8], int ks) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = ks * 32 + c * 4;
        if (live && vec && k + 4 <= K) {
          dst[c] = __ldg(reinterpret_cast<const float4*>(row + k));
        } else {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (live && k + j < K) ? __ldg(row + k + j) : 0.f;
          dst[c] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    load(cur, ks0);
    for (int it = 0; it < n_it; ++it) {
      if (it + 1 < n_it) load(nxt, ks0 + it + 1);
      wait_empty(it);
      uint8_t* dst = smem + (it % stages) * stage_bytes + t * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 x4 = cur[c];
        float4 v;
        v.x = round_tf32(x4.x); v.y = round_tf32(x4.y); v.z = round_tf32(x4.z); v.w = round_tf32(x4.w);
        *reinterpret_cast<float4*>(dst + (((c ^ (t & 7)) & 7) << 4)) = v;
        if (split) {
          float4 r;
          r.x = round_tf32(x4.x - v.x); r.y = round_tf32(x4.y - v.y); r.z = round_tf32(x4.z - v.z);
          r.w = round_tf32(x4.w - v.w);
          *reinterpret_cast<float4*>(dst + kBlockBytes + (((c ^ (t & 7)) & 7) << 4)) = r;
        }
      }
      arrive_full(it);
#pragma unroll
      for (int c = 0; c < 8; ++c) cur[c] = nxt[c];
    }
  }
};

}  // namespace tc
}  // namespace s2t
