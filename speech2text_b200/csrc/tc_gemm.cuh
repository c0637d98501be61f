// bf16 x bf16 -> fp32 contractions on the 5th-generation tensor cores.
//
//     C[m, n] (+)= sum_k A[m, k] * B[n, k]
//
// One CTA owns a 128 x BN output tile whose accumulator lives in TMEM; a single elected thread
// issues tcgen05.mma (M=128, N=BN, K=16) over 64-wide K blocks that arrive in shared memory
// through a kStages-deep mbarrier ring.  Warp roles (192 threads):
//   warp 0      bulk-copy issuer (cp.async.bulk, TMA engine) for the packed operands
//   warp 1      TMEM allocation + MMA issue + commits
//   warps 2..5  (a) optional on-the-fly A producer -- each thread builds one 128-byte row of the
//               A block (e.g. act(am + lm[ranges]) -> bf16) directly in the swizzled smem image,
//               so the operand never exists in HBM;  (b) epilogue: tcgen05.ld the accumulator
//               (one TMEM lane = one output row per thread) and hand 32-column chunks to the
//               epilogue functor.
// B is always a packed operand (tc_prims.cuh); A is packed (BulkA) or produced.
//
// Functor contracts
//   struct ASrc { static constexpr bool kBulk; ...
//       // kBulk:  const uint8_t* packed; int row_blocks;
//       // !kBulk: __device__ void produce(uint8_t* block, int m_tile, int kb, int t) const;  t in [0,128)
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
  int t;        // epilogue thread index, 0..127 (row inside the tile = quarter * 32 + lane)
  int row;      // row inside the 128-row tile
  int m_tile, n_tile, split;
  uint8_t* scratch;
  int scratch_bytes;
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct BulkA {
  static constexpr bool kBulk = true;
  const uint8_t* packed;
  int row_blocks;
  __device__ void produce(uint8_t*, int, int, int) const {}
};

constexpr int kGemmThreads = 192;

template <int BN, int kStages>
constexpr size_t gemm_stream_smem_bytes() {
  return (size_t)kStages * (kBlockBytes + (BN / 128) * kBlockBytes) + 1024 /*align*/ + 256 /*barriers*/;
}

template <int BN, int kStages, class ASrc, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_stream_kernel(ASrc asrc, const uint8_t* __restrict__ b_packed, int b_row_blocks, int k_blocks, int k_splits,
                   Epi epi) {
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
  constexpr int kABytes = kBlockBytes;
  constexpr int kBBytes = (BN / 128) * kBlockBytes;
  constexpr int kStageBytes = kABytes + kBBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int per = (k_blocks + k_splits - 1) / k_splits;
  const int kb0 = split * per;
  const int kb1 = min(k_blocks, kb0 + per);
  const int n_it = kb1 - kb0;
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
        const int s = it % kStages, ph = (it / kStages) & 1, kb = kb0 + it;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full[s], (ASrc::kBulk ? kABytes : 0) + kBBytes);
        if constexpr (ASrc::kBulk) {
          bulk_copy_g2s(sa, asrc.packed + packed_block_index(m_tile, kb, asrc.row_blocks) * kBlockBytes, kABytes,
                        &full[s]);
        }
        bulk_copy_g2s(sb, b_packed + packed_block_index(n_tile * (BN / 128), kb, b_row_blocks) * kBlockBytes,
                      kBBytes, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, BN);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
          umma_bf16(tmem_base, umma_smem_desc(sa + k4 * kUmmaK * 2), umma_smem_desc(sb + k4 * kUmmaK * 2), idesc,
                    (it > 0) || (k4 > 0));
        }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    const int t = (warp - 2) * 32 + lane;
    if constexpr (!ASrc::kBulk) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1, kb = kb0 + it;
        mbar_wait(&empty[s], ph ^ 1);
        asrc.produce(smem + s * kStageBytes, m_tile, kb, t);
        fence_proxy_async_smem();
        mbar_arrive(&full[s]);
      }
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

template <int BN, int kStages, class ASrc, class Epi>
int launch_gemm_stream(const ASrc& asrc, const uint8_t* b_packed, int b_row_blocks, int m_tiles, int n_tiles,
                       int k_blocks, int k_splits, const Epi& epi, cudaStream_t stream, const char* what) {
  if (m_tiles <= 0 || n_tiles <= 0 || k_blocks <= 0) return 0;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > k_blocks) k_splits = k_blocks;
  auto kern = gemm_stream_kernel<BN, kStages, ASrc, Epi>;
  constexpr size_t smem = gemm_stream_smem_bytes<BN, kStages>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString(e));
      return 2;
    }
    configured = true;
  }
  dim3 grid(n_tiles, m_tiles, k_splits);
  S2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", what);
  ProfScope prof(what, stream);
  kern<<<grid, kGemmThreads, smem, stream>>>(asrc, b_packed, b_row_blocks, k_blocks, k_splits, epi);
  return check_launch(what);
}

// ---- epilogues --------------------------------------------------------------------------------
// C[m * ldc + n] = acc (or += with atomics when partial sums from several splits / CTAs meet)
struct StoreRowMajorEpi {
  float* C;
  int64_t ldc;
  int M, N;
  bool atomic;
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
        else row[n + j] = acc[j];
      }
    }
  }
};

// ---- packing ----------------------------------------------------------------------------------
// fp32 src(r, k) = src[r * row_stride + k * col_stride]  ->  bf16 packed operand of row_blocks x k_blocks
// blocks, zero padded beyond (rows, K).
int pack_operand(const float* src, int64_t row_stride, int64_t col_stride, int rows, int K, int row_blocks,
                 int k_blocks, uint8_t* dst, cudaStream_t stream);

}  // namespace tc
}  // namespace s2t
