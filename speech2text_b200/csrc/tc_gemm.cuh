// bf16 x bf16 -> fp32 contractions on the 5th-generation tensor cores.
//
//     K-major  (kMn = false):  C[m, n] (+)= sum_k A[m, k]  * B[n, k]     operands stored (rows, K)
//     MN-major (kMn = true):   C[m, n] (+)= sum_k At[k, m] * Bt[k, n]    operands stored (K, rows)
//
// Persistent kernel, one CTA per SM, looping over 128 x BN output tiles.  The accumulators live in
// TMEM (two buffers); a single elected thread issues tcgen05.mma (M=128, N=BN, K=16) over 64-deep
// K steps that arrive in shared memory through a kStages-deep mbarrier ring which keeps running
// across tile boundaries.  Warp roles:
//   warp 0       bulk-copy issuer (cp.async.bulk, TMA engine) for the packed operands
//   warp 1       TMEM allocation + MMA issue + commits
//   warps 4..    epilogue groups of 4 warps: tcgen05.ld the finished accumulator (one TMEM lane = one
//                output row per thread) and hand 32-column chunks to the epilogue functor, while the
//                MMA warp is already filling the other TMEM buffer with the next tile; bulk-fed
//                kernels run two groups (half of the columns each), producer-fed kernels one
//   last 8 warps (only with an on-the-fly A source) producers: build the A stage, e.g.
//                act(am + lm[ranges]) -> bf16, directly in the swizzled smem image, so the operand
//                never exists in HBM
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
//       // !kBulk: __device__ void run(const ProdCtx&) const;  called by kProdThreads threads per tile; must,
//       //         for it in [0, n_it): pc.wait_empty(it); fill pc.stage(it); pc.arrive_full(it).
//   };
//   struct Epi {
//     struct State {...};                       // per-thread (= per output row) running state
//     __device__ void begin(State&, const EpiCtx&) const;
//     __device__ void chunk(State&, const EpiCtx&, int n, const float (&acc)[32]) const;   // 32 columns from n
//     __device__ void end(State&, const EpiCtx&) const;
//   };
// Epi::kScratchBytes of shared memory are reserved for CTA-level reductions of the epilogue
// per epilogue group (EpiCtx::scratch); epi_sync(ctx) is a barrier over the 128 threads of a group.
#pragma once
#include <utility>
#include <cstdio>
#include <type_traits>
#include <cstring>

#include "common.cuh"
#include "tc_prims.cuh"

namespace s2t {
namespace tc {

struct EpiCtx {
  int m;        // global output row of this thread
  int t;        // epilogue thread index, 0..127
  int row;      // row inside the 128-row tile
  int m_tile, n_tile, split, batch;
  int group;   // epilogue group: each group of 128 threads drains its own range of accumulator columns
  int col0;    // first output column of this group's range
  int ncols;   // columns in the range (BN / groups)
  int part;    // n_tile * groups + group: index of the (row, column-range) partial
  uint8_t* scratch;  // this group's Epi::kScratchBytes
  int scratch_bytes;
  int dbg;
  // Column sums (bias gradients), Epi::kColSums: when the CTA keeps ONE column tile for its whole life, the epilogue
  // adds its per-chunk column sums into this shared-memory row (index = column - tile_col0) and the kernel hands every
  // column's total to Epi::flush_colsum once, at the end -- instead of one global atomic per column, warp and tile,
  // all CTAs on the same few hundred addresses.  Null: add to global memory directly.
  float* colsum;
  int tile_col0;
};

// barrier over the 128 threads of one epilogue group
__device__ __forceinline__ void epi_sync(const EpiCtx& ctx) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + ctx.group) : "memory");
}

struct BulkA {
  static constexpr bool kBulk = true;
  const uint8_t* packed;
  int row_blocks;
};

// Warp roles of the persistent kernel.  Epilogue warps are 4..7 so that (warp & 3) is the TMEM lane
// quarter each may read; warps 2..3 only take part in the CTA-wide barriers.
constexpr int kCtrlWarps = 4;
constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kGroupBytes = 64 * 128;  // one MN-major group: 64 k-rows x 64 elements

#ifndef S2T_PAIR
#define S2T_PAIR 2
#endif
#ifndef S2T_BULK_EPI_GROUPS
#define S2T_BULK_EPI_GROUPS 4
#endif
// Epilogue groups: kernels fed purely by bulk copies have no producer warps and spend the thread
// budget on a second epilogue group (each group drains half of the accumulator columns).
// An epilogue functor may ask for fewer groups (static constexpr int kBulkGroups) when it needs the registers.
template <class Epi, class = void>
struct EpiBulkGroups {
  static constexpr int value = S2T_BULK_EPI_GROUPS;
};
template <class Epi>
struct EpiBulkGroups<Epi, decltype((void)Epi::kBulkGroups)> {
  static constexpr int value = Epi::kBulkGroups;
};
template <class Epi, class = void>
struct EpiColSums {
  static constexpr bool value = false;
};
template <class Epi>
struct EpiColSums<Epi, decltype((void)Epi::kColSums)> {
  static constexpr bool value = Epi::kColSums;
};
template <class ASrc, class Epi>
constexpr int gemm_epi_groups() {
  return ASrc::kBulk ? EpiBulkGroups<Epi>::value : 1;
}
template <class ASrc, class Epi>
constexpr int gemm_threads() {
  return 32 * (kCtrlWarps + kEpiWarps * gemm_epi_groups<ASrc, Epi>() + (ASrc::kBulk ? 0 : kProdWarps));
}

// kKind 3 = 3xF16: every fp32 operand is held as hi = f16(x) and lo = f16(x - hi) (two 64-column blocks per stage),
// D += A_lo B_hi + A_hi B_lo + A_hi B_hi: the same ~2^-21 accuracy as 3xTF32 at twice the tensor rate and half the
// operand bytes (operands must fit the f16 range: callers pre-scale by powers of two and undo it in the epilogue).
// kKind: 0 = bf16 operands; 1 = tf32 (fp32 in smem, single pass); 2 = 3xTF32: every fp32 operand is held
// as big = rn_tf32(x) and small = x - big, and D += A_big B_big + A_big B_small + A_small B_big, which
// recovers fp32-level accuracy (error ~2^-21) on the tensor cores.
// kKind 5 = 2xF16: as 3xF16 but the A operand is EXACT in f16 (bf16 activations inside the f16 range: 8 mantissa bits fit
// the 11 of a half) and carries no lo part: D += A B_lo + A B_hi -- two MMAs per product instead of three, half the A stage.
template <int BN, int kKind>
constexpr int gemm_stage_bytes() {
  return kKind == 5 ? kBlockBytes + 2 * (BN / 128) * kBlockBytes : (kKind >= 2 ? 2 : 1) * (kBlockBytes + (BN / 128) * kBlockBytes);
}
// kBRes > 0: the B operand of the CTA's column tile (kBRes k-steps) stays resident in shared memory and the ring
// carries A stages only
template <int BN, int kStages, int kKind, class ASrc, class Epi, int kBRes = 0>
constexpr size_t gemm_stream_smem_bytes_for(int stages) {
  return (size_t)stages * (kBRes > 0 ? kBlockBytes : gemm_stage_bytes<BN, kKind>()) +
         (size_t)kBRes * (BN / 128) * kBlockBytes +
         (size_t)gemm_epi_groups<ASrc, Epi>() * (((Epi::kScratchBytes + 127) / 128) * 128) + 1024 /*align*/ + 256 /*barriers*/ +
         (EpiColSums<Epi>::value ? BN * sizeof(float) : 0);
}
// ring depth actually used: the requested one, less when the epilogue scratch of all groups would not fit
template <int BN, int kStages, int kKind, class ASrc, class Epi, int kBRes = 0>
constexpr int gemm_eff_stages() {
  int s = kStages;
  while (s > 2 && gemm_stream_smem_bytes_for<BN, kStages, kKind, ASrc, Epi, kBRes>(s) > 227 * 1024) --s;
  return s;
}
template <int BN, int kStages, int kKind, class ASrc, class Epi, int kBRes = 0>
constexpr size_t gemm_stream_smem_bytes() {
  return gemm_stream_smem_bytes_for<BN, kStages, kKind, ASrc, Epi, kBRes>(
      gemm_eff_stages<BN, kStages, kKind, ASrc, Epi, kBRes>());
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
  // batched launches: operand offsets per batch, in 128-row blocks for K-major operands and in
  // 64-row k-steps for MN-major operands
  int a_batch_off = 0;
  int b_batch_off = 0;
  int dbg = 0;  // experiment flags (S2T_DBG): 1 skip epilogue functor, 2 only the big x big MMA, 16 no loads (MMA only)
  // S2T_TRACE=<kernel label>: CTA 0 records (clock, role event, index) words here; the launcher prints them
  unsigned long long* trace = nullptr;
  // Live row blocks, or null.  Padding frames fill whole 128-row blocks of the joiner lattice; live_idx lists the
  // blocks that hold at least one real frame (ascending), live_prefix[i] counts the live blocks before block i.  With
  // the list a row-wise (K-major) contraction walks the live row tiles of [live_off, live_off + m_tiles) only and a
  // reduction over the rows (MN-major, bulk-fed) walks the k-steps of the live blocks only, both evenly spread over
  // the CTAs / the k-splits (a static round-robin over all tiles leaves the makespan where it was), and nothing ever
  // reads or writes a dead block.  Single CTAs only (kCluster == 1).
  const int* live_idx = nullptr;
  const int* live_prefix = nullptr;
  int live_off = 0;
};

constexpr int kTraceCap = 4096;
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int code, int idx) {
  if (trace != nullptr && blockIdx.x == 0) {
    const unsigned long long slot = atomicAdd(trace, 1ull);
    if (slot < (unsigned long long)kTraceCap)
      trace[1 + slot] = ((unsigned long long)clock64() << 16) | ((unsigned long long)(code & 15) << 12) | (unsigned)(idx & 0xfff);
  }
}

// What an on-the-fly A producer sees for one tile: kProdThreads threads fill stage(it) for it in [0, n_it).
struct ProdCtx {
  int m_tile, n_tile, batch, ks0, n_it;
  bool valid;  // false: row tile past the end (odd CTA of a cluster): fill zeros, write no by-product
  int t;  // producer thread index, 0 .. kProdThreads-1
  uint8_t* smem;
  int stage_bytes, stages;
  uint32_t it0;  // ring position of the tile's first k-step
  uint64_t* full;
  uint64_t* empty;
  unsigned long long* trace = nullptr;
  __device__ __forceinline__ void mark(int code, int it) const {
    if (t == 0) trace_mark(trace, code, (int)(it0 + it));
  }
  __device__ __forceinline__ uint8_t* stage(int it) const { return smem + ((it0 + it) % stages) * stage_bytes; }
  __device__ __forceinline__ void wait_empty(int it) const {
    const uint32_t g = it0 + it;
    mbar_wait(&empty[g % stages], ((g / stages) & 1) ^ 1);
  }
  // every producer thread publishes its generic-proxy writes to the async proxy; one arrival per warp
  // (256 per-thread arrivals on one mbarrier serialise on the shared-memory atomic unit)
  __device__ __forceinline__ void arrive_full(int it) const {
    fence_proxy_async_smem();
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&full[(it0 + it) % stages]);
  }
};

struct TileCoord {
  int n_tile, m_tile, batch, split, ks0, n_it;
  bool valid;  // false: the odd CTA of a cluster past the last row tile (runs the protocol, produces nothing)
};

// Persistent, warp-specialised contraction kernel: one CTA per SM loops over output tiles
// (n fastest, so the CTAs running together share A rows in L2).  The smem operand ring and the
// two TMEM accumulator buffers run across tile boundaries: while the epilogue warps drain tile i
// from one TMEM buffer, the MMA warp already accumulates tile i+1 into the other and the copy /
// producer warps fill the ring for it.
// kCluster = 2 (CTA pair, tcgen05 cta_group::2): the two CTAs of a cluster own neighbouring row tiles of the same
// column tile and k range.  Each keeps its own 128 A rows and only HALF of the B rows of every stage; the leader
// CTA (rank 0) issues one M = 256 MMA that reads both shared memories and writes 128 accumulator lanes into each
// CTA's TMEM, so a stage pulls a third less through the L2 -> SM path that bounds the mainloop.  The peer forwards
// "my half of stage s has landed" to the leader's pfull barrier; the leader's commits release the stage (empty) and
// publish the accumulator (tfull) in both CTAs; both CTAs' epilogue warps hand the TMEM buffer back on the leader's
// tempty barrier.
// kBRes > 0 (B-stationary): a CTA keeps ONE column tile for its whole life, loads that tile's B operand (k_steps <= kBRes
// k-steps) into shared memory once and streams only A through the ring: a short-K contraction then pulls 16 KB instead
// of 48 KB per k-step through the L2 -> SM path, which is what bounds the streaming mainloop (profiles/README.md).
template <int BN, int kStagesReq, bool kMn, int kKind, class ASrc, class Epi, int kCluster = 1, int kBRes = 0>
__global__ void __launch_bounds__(gemm_threads<ASrc, Epi>(), 1)
gemm_stream_kernel(ASrc asrc, const uint8_t* __restrict__ b_packed, int b_row_blocks, int m_tiles, int n_tiles,
                   int batches, int k_steps, int k_splits, Epi epi, MnDebug mn) {
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
  static_assert(kBRes == 0 || (!kMn && kKind == 0 && ASrc::kBulk && kCluster == 1),
                "B-stationary mode: K-major bf16, bulk-fed, no cluster");
  constexpr int kParts = kKind >= 2 ? 2 : 1;
  constexpr int kABytes = (kKind == 5 ? 1 : kParts) * kBlockBytes;
  constexpr int kBPart = (BN / 128) * kBlockBytes / kCluster;  // this CTA's rows of one B part
  constexpr int kBBytes = kParts * kBPart;
  constexpr int kStageBytes = kBRes > 0 ? kABytes : kABytes + kBBytes;
  // a pair's stages are smaller (half of B): the ring the launcher sized for single CTAs holds more of them, which the
  // longer peer -> leader -> commit round trip of a stage needs
  constexpr int kStages1 = gemm_eff_stages<BN, kStagesReq, kKind, ASrc, Epi, kBRes>();
  constexpr int kStages = kCluster == 1 ? kStages1 : (kStages1 * gemm_stage_bytes<BN, kKind>()) / kStageBytes;
  constexpr int kGroups = gemm_epi_groups<ASrc, Epi>();
  constexpr int kScratch = ((Epi::kScratchBytes + 127) / 128) * 128;  // per epilogue group
  constexpr int kTmemCols = 2 * BN;  // two accumulator buffers
  static_assert(kKind < 2 || !ASrc::kBulk, "split operands expect an on-the-fly A producer that writes big|small");
  static_assert(kKind == 0 || !kMn, "tf32 / split operands are K-major only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* bres = smem + kStages * kStageBytes;  // resident B (kBRes k-steps), empty in streaming mode
  uint8_t* scratch = bres + kBRes * kBPart;
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + kGroups * kScratch);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres_full = tempty + 2;
  uint64_t* pfull = bres_full + 1;  // leader of a pair: the peer's half of stage s has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull + kStages);
  constexpr bool kColSums = EpiColSums<Epi>::value;
  float* colsum_smem = reinterpret_cast<float*>(tmem_slot + 4);  // [BN] when kColSums

  static_assert(kCluster == 1 || kCluster == 2, "clusters of one or two CTAs");
  static_assert(kCluster == 1 || !kMn, "CTA pairs are built for K-major operands");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = kCluster > 1 ? (int)cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], ASrc::kBulk ? 1 : 1 + kProdWarps);
      mbar_init(&empty[s], 1);
      mbar_init(&pfull[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      // one CTA: every epilogue thread arrives; pair: one lane per epilogue warp of both CTAs, on the leader's barrier
      mbar_init(&tempty[i], kCluster == 1 ? 32 * kEpiWarps * kGroups : kCluster * kEpiWarps * kGroups);
    }
    mbar_init(bres_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (kCluster == 1) tmem_alloc<kTmemCols>(tmem_slot);
    else tmem_alloc_pair<kTmemCols>(tmem_slot);
  }
  if constexpr (kColSums) {
    for (int i = threadIdx.x; i < BN; i += blockDim.x) colsum_smem[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel of the stream; nothing
  // below (the live-tile list included) is read before that kernel has completed
  griddep_wait();
  griddep_launch_dependents();
  // streaming: tiles of all (n, m, split, batch), n fastest; B-stationary: the CTA's own column tile, row tiles strided
  const int first_tile = kBRes > 0 ? (int)blockIdx.x / n_tiles : (int)blockIdx.x / kCluster;
  const int tile_stride = kBRes > 0 ? (int)gridDim.x / n_tiles : (int)gridDim.x / kCluster;
  // live-block list: lo = first list entry of this launch's block range, n_live = live blocks in it
  // (a CTA pair walks the list as well: its two CTAs take consecutive LIVE row tiles -- they share the column tile and the
  // k range, which row tiles they hold is free)
  const bool listed = (kCluster == 1 || !kMn) && (kMn ? ASrc::kBulk : true) && mn.live_idx != nullptr;
  const int live_lo = listed ? mn.live_prefix[mn.live_off] : 0;
  const int n_live = listed ? mn.live_prefix[mn.live_off + (kMn ? (k_steps + 1) / 2 : m_tiles)] - live_lo : 0;
  const int m_rows = (listed && !kMn) ? n_live : m_tiles;      // row tiles to walk
  const int k_live = (listed && kMn) ? 2 * n_live : k_steps;    // k-steps to walk (two per live 128-row block)
  const int m_groups = (m_rows + kCluster - 1) / kCluster;
  const int num_tiles = kBRes > 0 ? m_rows : n_tiles * m_groups * batches * k_splits;
  const int per = (k_live + k_splits - 1) / k_splits;
  // position in the walk -> row tile / k-step of the launch
  auto row_tile = [&](int j) { return (listed && !kMn) ? mn.live_idx[live_lo + j] - mn.live_off : j; };
  auto k_step = [&](int j) { return (listed && kMn) ? 2 * (mn.live_idx[live_lo + (j >> 1)] - mn.live_off) + (j & 1) : j; };
  constexpr uint16_t kCtaMask = (1u << kCluster) - 1;
  auto decode = [&](int tile) {
    TileCoord c;
    if constexpr (kBRes > 0) {
      c.n_tile = (int)blockIdx.x % n_tiles;
      c.m_tile = row_tile(tile);
      c.valid = true;
      c.split = 0;
      c.batch = 0;
      c.ks0 = 0;
      c.n_it = k_steps;
      return c;
    }
    c.n_tile = tile % n_tiles;
    int rest = tile / n_tiles;
    const int walk_m = (rest % m_groups) * kCluster + crank;
    c.valid = walk_m < m_rows;
    c.m_tile = c.valid ? row_tile(walk_m) : walk_m;
    rest /= m_groups;
    c.split = rest % k_splits;
    c.batch = rest / k_splits;
    c.ks0 = c.split * per;  // in walk positions: the bulk issuer maps them to k-steps
    c.n_it = min(k_live, c.ks0 + per) - c.ks0;
    return c;
  };


  if (warp == 0) {
    // ---------------- bulk-copy issuer ----------------
    if (lane == 0 && !(mn.dbg & 16)) {
      uint32_t git = 0;
      if constexpr (kBRes > 0) {
        if (first_tile < num_tiles) {
          mbar_arrive_expect_tx(bres_full, (uint32_t)k_steps * kBPart);
          const int nt = (int)blockIdx.x % n_tiles;
          for (int ks = 0; ks < k_steps; ++ks)
            bulk_copy_g2s(bres + ks * kBPart,
                          b_packed + packed_block_index(nt * (BN / 128), ks, b_row_blocks) * kBlockBytes, kBPart, bres_full);
        }
      }
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const TileCoord c = decode(tile);
        for (int it = 0; it < c.n_it; ++it) {
          const int ks = k_step(c.ks0 + it);
          const int s = git % kStages;
          const uint32_t g_now = git++;
          mbar_wait(&empty[s], ((g_now / kStages) & 1) ^ 1);
          uint8_t* sa = smem + s * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if constexpr (kBRes > 0) {
            mbar_arrive_expect_tx(&full[s], kABytes);
            bulk_copy_g2s(sa, asrc.packed + packed_block_index(c.m_tile, ks, asrc.row_blocks) * kBlockBytes, kABytes, &full[s]);
            continue;
          }
          mbar_arrive_expect_tx(&full[s], (ASrc::kBulk ? kABytes : 0) + kBBytes);
          if constexpr (!kMn) {
            if constexpr (ASrc::kBulk) {
              const int mt = c.valid ? c.m_tile : m_tiles - 1;  // keep the byte count; the result is discarded
              bulk_copy_g2s(sa, asrc.packed + packed_block_index(mt + c.batch * mn.a_batch_off, ks, asrc.row_blocks) * kBlockBytes,
                            kABytes, &full[s]);
            }
            const size_t boff =
                packed_block_index(c.n_tile * (BN / 128) + c.batch * mn.b_batch_off, ks, b_row_blocks) * kBlockBytes;
            // a pair: this CTA's half of the column tile's rows (contiguous in the packed layout)
            const size_t roff = boff + (size_t)crank * kBPart;
            bulk_copy_g2s(sb, b_packed + roff, kBPart, &full[s]);
            if constexpr (kKind >= 2) bulk_copy_g2s(sb + kBPart, mn.b_small + roff, kBPart, &full[s]);
            trace_mark(mn.trace, 7, (int)git);
          } else {
            // k-step = 64 contraction rows = half of a 128-row block; group = 64 columns = one column block
            if constexpr (ASrc::kBulk) {
              const int ka = ks + c.batch * mn.a_batch_off;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                bulk_copy_g2s(sa + g * kGroupBytes,
                              asrc.packed + packed_block_index(ka >> 1, c.m_tile * 2 + g, asrc.row_blocks) * kBlockBytes +
                                  (size_t)(ka & 1) * kGroupBytes,
                              kGroupBytes, &full[s]);
              }
            }
            const int kb2 = ks + c.batch * mn.b_batch_off;
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) {
              bulk_copy_g2s(sb + g * kGroupBytes,
                            b_packed + packed_block_index(kb2 >> 1, c.n_tile * (BN / 64) + g, b_row_blocks) * kBlockBytes +
                                (size_t)(kb2 & 1) * kGroupBytes,
                            kGroupBytes, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && kCluster > 1 && crank != 0) {
      // peer of a pair: no MMA of its own; tells the leader when its half of each stage is in shared memory
      uint32_t git = 0;
      for (int tile = first_tile; tile < num_tiles && !(mn.dbg & 16); tile += tile_stride) {
        const TileCoord c = decode(tile);
        for (int it = 0; it < c.n_it; ++it, ++git) {
          const int s = git % kStages;
          mbar_wait(&full[s], (git / kStages) & 1);
          mbar_arrive_cluster(&pfull[s], 0);
        }
      }
    } else if (lane == 0) {
      constexpr int kM = 128 * kCluster;
      uint32_t idesc = kKind == 0 ? umma_idesc_bf16(kM, BN)
                                  : ((kKind == 3 || kKind == 5) ? umma_idesc_f16(kM, BN) : umma_idesc_tf32(kM, BN));
      if (kMn) idesc |= (1u << 15) | (1u << 16);  // A and B are MN-major
      auto mma16 = [](uint32_t acc, uint64_t da, uint64_t db, uint32_t id, bool accum) {
        if constexpr (kCluster == 1) umma_bf16(acc, da, db, id, accum);
        else umma_f16_pair(acc, da, db, id, accum);
      };
      auto mma32 = [](uint32_t acc, uint64_t da, uint64_t db, uint32_t id, bool accum) {
        if constexpr (kCluster == 1) umma_tf32(acc, da, db, id, accum);
        else umma_tf32_pair(acc, da, db, id, accum);
      };
      uint32_t git = 0, lt = 0;
      if constexpr (kBRes > 0) {
        if (first_tile < num_tiles) mbar_wait(bres_full, 0);
      }
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const TileCoord c = decode(tile);
        if (c.n_it <= 0) continue;
        const uint32_t buf = lt & 1;
        // epilogue (of both CTAs of a pair) has drained this accumulator
        if constexpr (kCluster == 1) mbar_wait(&tempty[buf], ((lt >> 1) & 1) ^ 1);
        else mbar_wait_cluster(&tempty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        bool first = true;  // the first k-step actually issued overwrites the accumulator
        for (int it = 0; it < c.n_it; ++it) {
          const int s = git % kStages;
          const uint32_t g_now = git++;
          if (!(mn.dbg & 16)) {
            mbar_wait(&full[s], (g_now / kStages) & 1);
            if constexpr (kCluster > 1) mbar_wait_cluster(&pfull[s], (g_now / kStages) & 1);
          }
          trace_mark(mn.trace, 3, (int)g_now);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint32_t sb = kBRes > 0 ? smem_u32(bres + (c.ks0 + it) * kBPart) : sa + kABytes;
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
            const bool accum = !first || (k4 > 0);
            if constexpr (kKind == 0) {
              mma16(acc, da, db, idesc, accum);
            } else if constexpr (kKind == 1) {
              mma32(acc, da, db, idesc, accum);
            } else if constexpr (kKind == 3) {
              const uint64_t da_s = umma_smem_desc(sa + kBlockBytes + k4 * kUmmaK * 2);
              const uint64_t db_s = umma_smem_desc(sb + kBPart + k4 * kUmmaK * 2);
              mma16(acc, da_s, db, idesc, accum);  // small terms first
              mma16(acc, da, db_s, idesc, true);
              mma16(acc, da, db, idesc, true);
            } else if constexpr (kKind == 5) {
              const uint64_t db_s = umma_smem_desc(sb + kBPart + k4 * kUmmaK * 2);
              mma16(acc, da, db_s, idesc, accum);  // small term first
              mma16(acc, da, db, idesc, true);
            } else {
              const uint64_t da_s = umma_smem_desc(sa + kBlockBytes + k4 * kUmmaK * 2);
              const uint64_t db_s = umma_smem_desc(sb + kBPart + k4 * kUmmaK * 2);
              if (!(mn.dbg & 2)) {
                mma32(acc, da_s, db, idesc, accum);  // small terms first
                mma32(acc, da, db_s, idesc, true);
                mma32(acc, da, db, idesc, true);
              } else {
                mma32(acc, da, db, idesc, accum);
              }
            }
          }
          first = false;
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if constexpr (kCluster == 1) umma_commit(&empty[s]);
          else umma_commit_pair(&empty[s], kCtaMask);
        }
        if constexpr (kCluster == 1) umma_commit(&tfull[buf]);
        else umma_commit_pair(&tfull[buf], kCtaMask);
        trace_mark(mn.trace, 4, (int)lt);
        ++lt;
      }
    }
  } else if (warp >= kCtrlWarps && warp < kCtrlWarps + kEpiWarps * kGroups) {
    // ---------------- epilogue ----------------
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    const int group = (warp - kCtrlWarps) / kEpiWarps;
    constexpr int kCols = BN / kGroups;
    uint32_t lt = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
      const TileCoord c = decode(tile);
      if (c.n_it <= 0) continue;
      const uint32_t buf = lt & 1;
      mbar_wait(&tfull[buf], (lt >> 1) & 1);
      if (warp == kCtrlWarps && lane == 0) trace_mark(mn.trace, 5, (int)lt);
      tc_fence_after();
      EpiCtx ctx;
      ctx.row = quarter * 32 + lane;
      ctx.m = c.m_tile * 128 + ctx.row;
      ctx.t = ((warp - kCtrlWarps) % kEpiWarps) * 32 + lane;
      ctx.m_tile = c.m_tile;
      ctx.n_tile = c.n_tile;
      ctx.split = c.split;
      ctx.batch = c.batch;
      ctx.group = group;
      ctx.col0 = c.n_tile * BN + group * kCols;
      ctx.ncols = kCols;
      ctx.part = c.n_tile * kGroups + group;
      ctx.scratch = scratch + group * kScratch;
      ctx.scratch_bytes = kScratch;
      ctx.dbg = mn.dbg;
      ctx.colsum = nullptr;
      ctx.tile_col0 = c.n_tile * BN;
      if constexpr (kColSums) {
        // one column tile per CTA: B-stationary launches, or a streaming launch with a single column tile and no splits
        if (kBRes > 0 || (n_tiles == 1 && k_splits == 1 && batches == 1)) ctx.colsum = colsum_smem;
      }
      if (c.valid) {
        typename Epi::State st;
        epi.begin(st, ctx);
#pragma unroll 1
        for (int cc = 0; cc < kCols / 32; ++cc) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + group * kCols + cc * 32, v);
          if (!(mn.dbg & 1)) epi.chunk(st, ctx, ctx.col0 + cc * 32, v);
        }
        epi.end(st, ctx);
      }
      if (warp == kCtrlWarps && lane == 0) trace_mark(mn.trace, 6, (int)lt);
      tc_fence_before();
      if constexpr (kCluster == 1) {
        mbar_arrive(&tempty[buf]);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty[buf], 0);
      }
      ++lt;
    }
    if constexpr (kColSums) {
      if (kBRes > 0 || (n_tiles == 1 && k_splits == 1 && batches == 1)) {
        // all epilogue warps have added their sums: hand the CTA's totals over, one global atomic per column
        constexpr int kEpiThreads = 32 * kEpiWarps * kGroups;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + kGroups), "n"(kEpiThreads) : "memory");
        const int col0 = (kBRes > 0 ? (int)blockIdx.x % n_tiles : 0) * BN;
        for (int i = (int)threadIdx.x - 32 * kCtrlWarps; i < BN; i += kEpiThreads) {
          const float v = colsum_smem[i];
          if (v != 0.f) epi.flush_colsum(col0 + i, v);
        }
      }
    }
  } else if (warp >= kCtrlWarps + kEpiWarps * kGroups) {
    // ---------------- on-the-fly A producers ----------------
    if constexpr (!ASrc::kBulk) {
      uint32_t git = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const TileCoord c = decode(tile);
        if (c.n_it <= 0) continue;
        ProdCtx pc;
        pc.m_tile = c.m_tile;
        pc.valid = c.valid;
        pc.n_tile = c.n_tile;
        pc.batch = c.batch;
        pc.ks0 = c.ks0;
        pc.n_it = c.n_it;
        pc.t = (warp - kCtrlWarps - kEpiWarps * kGroups) * 32 + lane;
        pc.smem = smem;
        pc.stage_bytes = kStageBytes;
        pc.stages = kStages;
        pc.it0 = git;
        pc.full = full;
        pc.empty = empty;
        pc.trace = mn.trace;
        asrc.run(pc);
        git += c.n_it;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();  // nobody leaves while the peer may still signal or multicast
  if (warp == 1) {
    if constexpr (kCluster == 1) tmem_dealloc<kTmemCols>(tmem_base);
    else tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

inline int gemm_sm_count() { return device_info().sms; }

// cudaFuncSetAttribute(max dynamic shared memory) once per kernel instantiation AND device
template <class Kern>
inline int configure_smem_once(Kern kern, size_t smem, bool (&done)[kMaxDevices], const char* what) {
  const int dev = device_info().device;
  if (!done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem, cudaGetErrorString(e));
      return 2;
    }
    done[dev] = true;
  }
  return 0;
}

// S2T_TRACE=<kernel label>: the third and fourth launch of that kernel record CTA 0's role timeline, printed to stderr
// when the launch has finished (developer aid; synchronises the stream).
struct TraceScope {
  unsigned long long* buf = nullptr;
  cudaStream_t st;
  const char* what;
  TraceScope(const char* what_, cudaStream_t stream, MnDebug& mn) : st(stream), what(what_) {
    static unsigned long long* trace_buf = nullptr;
    static int trace_left = 2;
    static int trace_seen = 0;  // the first launches of a process are cold
    const char* trace_env = getenv("S2T_TRACE");
    const bool tracing = trace_env && strstr(what, trace_env) && trace_seen++ >= 2 && trace_left > 0;
    if (!tracing) return;
    if (!trace_buf) cudaMalloc(&trace_buf, (kTraceCap + 1) * sizeof(unsigned long long));
    cudaMemsetAsync(trace_buf, 0, (kTraceCap + 1) * sizeof(unsigned long long), stream);
    mn.trace = trace_buf;
    buf = trace_buf;
    --trace_left;
  }
  ~TraceScope() {
    if (!buf) return;
    cudaStreamSynchronize(st);
    static unsigned long long host[kTraceCap + 1];
    cudaMemcpy(host, buf, sizeof(host), cudaMemcpyDeviceToHost);
    const unsigned long long n = host[0] < (unsigned long long)kTraceCap ? host[0] : kTraceCap;
    unsigned long long t0 = ~0ull;
    for (unsigned long long i = 0; i < n; ++i) if ((host[1 + i] >> 16) < t0) t0 = host[1 + i] >> 16;
    fprintf(stderr, "S2T_TRACE %s events=%llu\n", what, n);
    for (unsigned long long i = 0; i < n; ++i)
      fprintf(stderr, "T %llu %d %d\n", (host[1 + i] >> 16) - t0, (int)((host[1 + i] >> 12) & 15), (int)(host[1 + i] & 0xfff));
  }
};

// Launch with the programmatic-dependent-launch attribute (and the cluster shape, if any): the kernel may become resident
// while its predecessor in the stream drains; it calls griddep_wait() before it touches global memory.
// S2T_B200_NO_PDL=1 launches plainly (A/B switch; the device-side calls are no-ops then).
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("S2T_B200_NO_PDL") ? 0 : 1;
  return on == 1;
}

template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, int cluster,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

template <int BN, int kStages, bool kMn, int kKind, int kCluster = 1, class ASrc, class Epi>
int launch_gemm_stream(const ASrc& asrc, const uint8_t* b_packed, int b_row_blocks, int m_tiles, int n_tiles,
                       int k_steps, int k_splits, const Epi& epi, cudaStream_t stream, const char* what,
                       MnDebug mn = MnDebug(), int batches = 1) {
  if (m_tiles <= 0 || n_tiles <= 0 || k_steps <= 0 || batches <= 0) return 0;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > k_steps) k_splits = k_steps;
  auto kern = gemm_stream_kernel<BN, kStages, kMn, kKind, ASrc, Epi, kCluster>;
  constexpr size_t smem = gemm_stream_smem_bytes<BN, kStages, kKind, ASrc, Epi>();
  static_assert(smem <= 227 * 1024, "stage ring + epilogue scratch exceed the 227 KB of one CTA");
  static bool configured[kMaxDevices] = {};
  if (int rc = configure_smem_once(kern, smem, configured, what)) return rc;
  const long long tiles = (long long)((m_tiles + kCluster - 1) / kCluster) * n_tiles * batches * k_splits;  // per cluster
  const int max_clusters = gemm_sm_count() / kCluster;
  const int grid = kCluster * (int)(tiles < max_clusters ? tiles : max_clusters);
  {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("S2T_DBG") ? atoi(getenv("S2T_DBG")) : 0;
    mn.dbg = dbg;
  }
  TraceScope trace_scope(what, stream, mn);
  ProfScope prof(what, stream);
  {
    cudaError_t e = launch_pdl(kern, (unsigned)grid, (unsigned)gemm_threads<ASrc, Epi>(), smem, stream, kCluster, asrc, b_packed,
                               b_row_blocks, m_tiles, n_tiles, batches, k_steps, k_splits, epi, mn);
    if (e != cudaSuccess) {
      set_error("%s: launch: %s", what, cudaGetErrorString(e));
      return 2;
    }
  }
  return check_launch(what);
}

// B-stationary launch of a K-major bf16 bulk-fed contraction with k_steps <= kBRes (e.g. K = inner dim = 256: four
// k-steps, 128 KB of resident B per 256-column tile); anything else goes to the streaming kernel.
template <int BN, int kStages, int kBRes, class ASrc, class Epi>
int launch_gemm_bstationary(const ASrc& asrc, const uint8_t* b_packed, int b_row_blocks, int m_tiles, int n_tiles,
                            int k_steps, const Epi& epi, cudaStream_t stream, const char* what,
                            const MnDebug& live = MnDebug()) {
  constexpr size_t smem = gemm_stream_smem_bytes<BN, kStages, 0, ASrc, Epi, kBRes>();
  static_assert(smem <= 227 * 1024, "resident B + A ring + epilogue scratch exceed the 227 KB of one CTA");
  const int sms = gemm_sm_count();
  if (k_steps > kBRes || n_tiles > sms || getenv("S2T_B200_NO_BSTATIONARY")) {
    return launch_gemm_stream<BN, kStages, false, 0>(asrc, b_packed, b_row_blocks, m_tiles, n_tiles, k_steps, 1, epi, stream, what,
                                                     live);
  }
  if (m_tiles <= 0 || n_tiles <= 0 || k_steps <= 0) return 0;
  auto kern = gemm_stream_kernel<BN, kStages, false, 0, ASrc, Epi, 1, kBRes>;
  static bool configured[kMaxDevices] = {};
  if (int rc = configure_smem_once(kern, smem, configured, what)) return rc;
  // every CTA owns one column tile: grid = column tiles x (CTAs per column tile)
  int per_tile = sms / n_tiles;
  if (per_tile > m_tiles) per_tile = m_tiles;
  const int grid = per_tile * n_tiles;
  MnDebug mn = live;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("S2T_DBG") ? atoi(getenv("S2T_DBG")) : 0;
    mn.dbg = dbg;
  }
  TraceScope trace_scope(what, stream, mn);
  ProfScope prof(what, stream);
  {
    cudaError_t e = launch_pdl(kern, (unsigned)grid, (unsigned)gemm_threads<ASrc, Epi>(), smem, stream, 1, asrc, b_packed,
                               b_row_blocks, m_tiles, n_tiles, 1, k_steps, 1, epi, mn);
    if (e != cudaSuccess) {
      set_error("%s: launch: %s", what, cudaGetErrorString(e));
      return 2;
    }
  }
  return check_launch(what);
}

// ---- epilogues --------------------------------------------------------------------------------
// C[m * ldc + n] = acc (or += with atomics when partial sums from several splits / CTAs meet).
// An epilogue thread owns one output ROW (one TMEM lane), so storing its 32 columns directly would make
// every store instruction of the warp touch 32 different lines.  Each warp instead transposes 32 x 16
// sub-tiles through a padded shared-memory tile (row stride 20 words: conflict-free for the 16-byte
// writes by row and the 16-byte reads by 8 rows x 4 column groups) and writes 64 contiguous bytes per row.
constexpr int kTransposeRowWords = 20;
constexpr int kTransposeScratchBytes = 4 * 32 * kTransposeRowWords * 4;  // per epilogue group

// Hands the warp's 32 x 32 accumulator chunk (thread = row) back as f(r, c, v): v = columns c .. c+3 of the
// warp's row r, with the eight lanes l, l+8, l+16, l+24 ... of a call covering 64 contiguous bytes of a row.
// f is called 8 times per thread; `cols` (uniform over the warp) limits the columns that are needed.
template <class F>
__device__ __forceinline__ void warp_transposed_chunk(const EpiCtx& ctx, const float (&acc)[32], int cols, F&& f) {
  float* tile = reinterpret_cast<float*>(ctx.scratch) + (ctx.t >> 5) * (32 * kTransposeRowWords);
  const int lane = ctx.t & 31;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h * 16 >= cols) break;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(tile + lane * kTransposeRowWords + q * 4) =
          make_float4(acc[h * 16 + q * 4 + 0], acc[h * 16 + q * 4 + 1], acc[h * 16 + q * 4 + 2], acc[h * 16 + q * 4 + 3]);
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int r = pass * 8 + (lane & 7), cg = lane >> 3;
      const float4 v = *reinterpret_cast<const float4*>(tile + r * kTransposeRowWords + cg * 4);
      f(r, h * 16 + cg * 4, v);
    }
    __syncwarp();
  }
}

// Same idea for 2-byte outputs: the thread's 32 columns are 64 bytes (four 16-byte pieces).  f(r, q, v) hands
// piece q (columns 8 q .. 8 q + 7) of the warp's row r back; the four lanes l, l+8, l+16, l+24 of a call cover the
// 64 contiguous bytes of one row, so a warp-wide 16-byte store writes whole 32-byte sectors of eight rows.
template <class F>
__device__ __forceinline__ void warp_transposed_chunk_b16(const EpiCtx& ctx, const uint4 (&mine)[4], F&& f) {
  uint32_t* tile = reinterpret_cast<uint32_t*>(ctx.scratch) + (ctx.t >> 5) * (32 * kTransposeRowWords);
  const int lane = ctx.t & 31;
#pragma unroll
  for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(tile + lane * kTransposeRowWords + q * 4) = mine[q];
  __syncwarp();
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = pass * 8 + (lane & 7), q = lane >> 3;
    f(r, q, *reinterpret_cast<const uint4*>(tile + r * kTransposeRowWords + q * 4));
  }
  __syncwarp();
}

__device__ __forceinline__ void pack_row32_bf16(const float (&x)[32], uint4 (&out)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    out[c] = make_uint4(pack_bf16x2(x[c * 8 + 0], x[c * 8 + 1]), pack_bf16x2(x[c * 8 + 2], x[c * 8 + 3]),
                        pack_bf16x2(x[c * 8 + 4], x[c * 8 + 5]), pack_bf16x2(x[c * 8 + 6], x[c * 8 + 7]));
}

// max(*addr, v) for floats through the integer atomics (IEEE order: non-negative floats compare like ints, negative
// ones like reversed unsigneds); *addr must start at -inf or any float
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct StoreRowMajorEpi {
  float* C;
  int64_t ldc;
  int M, N;
  bool atomic;
  const float* bias = nullptr;  // added per column when not atomic
  float scale = 1.f;            // applied to the accumulator first (undoes a power-of-two operand pre-scale)
  // By-product: row_max[m] = max_n C[m, n], the stabiliser the simple-loss normaliser needs for every
  // projected row -- formed from the values on their way to memory instead of by a second pass over C.  Must hold
  // -inf on entry (column tiles meet through an atomic max).
  float* row_max = nullptr;
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  // The lane's bias values for the NEXT 32-column chunk (two 16-column halves x 4 columns) are fetched while the
  // current chunk is transposed: a bias load inside the store callback exposes one L2 round trip per pass, because
  // the producers' streaming loads leave nothing of the small L1 carve-out.
  struct State {
    float4 nb[2];
    float rm[4];  // running maxima of the rows pass * 8 + lane % 8 this lane serves after the transpose
  };
  __device__ __forceinline__ void load_bias(float4 (&b)[2], int n, int lane) const {
    const bool vec = (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n + h * 16 + (lane >> 3) * 4;
      if (bias == nullptr || atomic || col >= N) {
        b[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (vec && col + 4 <= N) {
        b[h] = __ldg(reinterpret_cast<const float4*>(bias + col));
      } else {
        b[h].x = __ldg(bias + col);
        b[h].y = col + 1 < N ? __ldg(bias + col + 1) : 0.f;
        b[h].z = col + 2 < N ? __ldg(bias + col + 2) : 0.f;
        b[h].w = col + 3 < N ? __ldg(bias + col + 3) : 0.f;
      }
    }
  }
  __device__ void begin(State& st, const EpiCtx& ctx) const {
    load_bias(st.nb, ctx.col0, ctx.t & 31);
    st.rm[0] = st.rm[1] = st.rm[2] = st.rm[3] = kNegInf;
  }
  __device__ void end(State& st, const EpiCtx& ctx) const {
    if (row_max == nullptr) return;
    const int lane = ctx.t & 31, m0 = ctx.m - lane;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      float v = st.rm[pass];  // the four lanes l, l+8, l+16, l+24 hold the four column groups of one row
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
      const int m = m0 + pass * 8 + (lane & 7);
      if (lane < 8 && m < M && v > kNegInf) atomic_max_float(row_max + m, v);
    }
  }
  __device__ void chunk(State& st, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    const int m0 = ctx.m - (ctx.t & 31);  // first row of this warp
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const float4 cur[2] = {st.nb[0], st.nb[1]};
    if (n + 32 < ctx.col0 + ctx.ncols) load_bias(st.nb, n + 32, ctx.t & 31);
    warp_transposed_chunk(ctx, acc, N - n, [&](int r, int c, float4 v) {
      const int m = m0 + r, col = n + c;
      if (m >= M || col >= N) return;
      const float4 b = cur[c >> 4];
      v.x = v.x * scale + b.x; v.y = v.y * scale + b.y; v.z = v.z * scale + b.z; v.w = v.w * scale + b.w;
      if (row_max != nullptr) {
        float mx = v.x;
        if (col + 1 < N) mx = fmaxf(mx, v.y);
        if (col + 2 < N) mx = fmaxf(mx, v.z);
        if (col + 3 < N) mx = fmaxf(mx, v.w);
        st.rm[r >> 3] = fmaxf(st.rm[r >> 3], mx);
      }
      float* dst = C + (int64_t)m * ldc + col;
      if (vec_ok && col + 4 <= N) {
        if (atomic) atomicAdd(reinterpret_cast<float4*>(dst), v);
        else *reinterpret_cast<float4*>(dst) = v;
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col + j < N) {
            if (atomic) atomicAdd(dst + j, e[j]);
            else dst[j] = e[j];
          }
        }
      }
    });
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

// Row-major fp32 (rows, K) -> bf16 packed operand of src (+ src2 when not null); col_sum (may be null) += its column sums.
int pack_rows_colsum(const float* src, const float* src2, int64_t ld, int rows, int K, int row_blocks, int k_blocks,
                     uint8_t* dst, float* col_sum, cudaStream_t stream);

// Several pack jobs in one launch (weights of a module).  kind 0: bf16 (k_blocks of 64); kind 1 / 2: the
// big / residual tf32 part of an fp32 operand (k_blocks of 32); kind 3 / 4: the hi / lo f16 part of scale * x
// (k_blocks of 64).  src(r, k) = src[r * row_stride + k * col_stride].
struct PackJob {
  const float* src;
  int64_t row_stride, col_stride;
  int rows, K, row_blocks, k_blocks;
  uint8_t* dst;
  int kind;
  float scale = 1.f;
};
struct PackJobs {
  static constexpr int kMax = 6;
  PackJob job[kMax];
  int64_t first[kMax + 1];  // first chunk (thread) of every job
  int n;
  float* fill = nullptr;  // optional: fill[0 .. fill_n) = fill_value by the threads behind the last job (saves a launch)
  int64_t fill_n = 0;
  float fill_value = 0.f;
};
int pack_jobs(const PackJob* list, int n, cudaStream_t stream, float* fill = nullptr, int64_t fill_n = 0, float fill_value = 0.f);

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
// dst_bf16 (optional): the same values once more as a bf16 operand of bf16_k_blocks 64-column blocks; columns the
// fp32 operand does not cover must already be zero there
int pack_f32_split(const PackSpec& p, uint8_t* dst_big, uint8_t* dst_small, cudaStream_t stream,
                   uint8_t* dst_bf16 = nullptr, int bf16_k_blocks = 0);
// hi / lo f16 halves of scale * f(x) (k_blocks of 64) and, optionally, bf16 f(x) in the same block geometry
int pack_f16_split(const PackSpec& p, float scale, uint8_t* dst_hi, uint8_t* dst_lo, uint8_t* dst_bf16,
                   cudaStream_t stream);

// On-the-fly K-major A for tf32: copies 128 rows x 32 fp32 of a row-major matrix into the swizzled stage.
// Eight lanes cover the 128 bytes one row contributes to a k-step, so a warp-wide load instruction reads
// four whole 128-byte lines; thread (warp w, lane l) serves rows 4w + l/8 + 32 i, i = 0..3.
struct RowCopyProducerF32 {
  static constexpr bool kBulk = false;
  const float* x;
  int64_t ld;
  int64_t M;
  int K;
  bool split;  // also write the residual block right after the big block (3xTF32)
  // optional by-product: the bf16 packed image of x (row_blocks = m tiles, 64-column blocks), written while
  // the first column tile of every row tile streams by -- the weight-gradient contraction reads it later
  uint8_t* bf16_pack = nullptr;
  int pack_row_blocks = 0;
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int c = lane & 7, rbase = warp * 4 + (lane >> 3);
    const bool emit = bf16_pack != nullptr && pc.n_tile == 0 && pc.valid;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const float* rowp[4];
    bool live[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = (int64_t)pc.m_tile * 128 + rbase + 32 * i;
      live[i] = m < M;
      rowp[i] = x + (live[i] ? m * ld : 0);
    }
    // three register buffers: the loads of k-steps it+1 and it+2 are in flight while step it is converted
    float4 buf[3][4];
    auto load = [&](float4 (&dst)[4], int ks) {
      const int k = ks * 32 + c * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (live[i] && vec && k + 4 <= K) {
          dst[i] = __ldg(reinterpret_cast<const float4*>(rowp[i] + k));
        } else {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (live[i] && k + j < K) ? __ldg(rowp[i] + k + j) : 0.f;
          dst[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    const int off = ((c ^ (rbase & 7)) & 7) << 4;
    auto emit_stage = [&](const float4 (&cur)[4], int it) {
      pc.wait_empty(it);
      uint8_t* dst = pc.stage(it) + rbase * 128 + off;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x4 = cur[i];
        float4 v;
        v.x = round_tf32(x4.x); v.y = round_tf32(x4.y); v.z = round_tf32(x4.z); v.w = round_tf32(x4.w);
        *reinterpret_cast<float4*>(dst + i * 32 * 128) = v;
        if (split) {
          float4 q;
          q.x = round_tf32(x4.x - v.x); q.y = round_tf32(x4.y - v.y); q.z = round_tf32(x4.z - v.z);
          q.w = round_tf32(x4.w - v.w);
          *reinterpret_cast<float4*>(dst + kBlockBytes + i * 32 * 128) = q;
        }
        if (emit) {
          const int ks = pc.ks0 + it;
          uint8_t* blk = bf16_pack + packed_block_index(pc.m_tile, ks >> 1, pack_row_blocks) * kBlockBytes;
          *reinterpret_cast<uint2*>(blk + block_chunk_offset(rbase + 32 * i, (ks & 1) * 4 + (c >> 1)) + (c & 1) * 8) =
              make_uint2(pack_bf16x2(x4.x, x4.y), pack_bf16x2(x4.z, x4.w));
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

// On-the-fly K-major A for the 3xF16 contraction: 128 rows x 64 fp32 of a row-major matrix -> hi | lo half blocks of
// the stage.  Sixteen lanes cover the 256 bytes one row contributes to a k-step; thread (warp w, lane l) serves rows
// 2w + l/16 + 16 i, i = 0..7.  Optional by-product: the bf16 packed image of x (same block geometry).
// TX = float or __nv_bfloat16 (BASELINE config 3's "bf16 joiner": the activations arrive as bf16 and are read as such --
// half the bytes, and the packed bf16 image kept for the weight gradient is then exact).
template <typename TX, bool kLo = true>
struct RowSplitProducerF16T {
  static constexpr bool kBulk = false;
  const TX* x;
  int64_t ld;
  int64_t M;
  int K;
  uint8_t* bf16_pack = nullptr;
  int pack_row_blocks = 0;
  // four consecutive elements as loaded: the conversion to fp32 happens when the step is emitted, NOT at the load --
  // a conversion next to the load would make the warp wait for the data before the barrier wait and the stores of the
  // step in front, i.e. take the loads out of flight
  using Raw = typename std::conditional<sizeof(TX) == 4, float4, uint2>::type;
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[4]) {
    if constexpr (sizeof(TX) == 4) {
      v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    } else {
      v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
      v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
  }
  __device__ void run(const ProdCtx& pc) const {
    const int warp = pc.t >> 5, lane = pc.t & 31;
    const int c = lane & 15, rbase = warp * 2 + (lane >> 4);
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool emit = bf16_pack != nullptr && pc.n_tile == 0 && pc.valid;
    const int off = rbase * 128 + ((((c >> 1) ^ (rbase & 7)) & 7) << 4) + (c & 1) * 8;
    // Two whole k-steps of raw rows (2 x 8 pieces per lane) are kept in flight: with one k-step every iteration costs a
    // full DRAM round trip, Little's law then caps the CTA at ~32 KB per microsecond.
    auto load = [&](Raw (&dst)[8], int ks) {
      const int k = ks * 64 + c * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t m = (int64_t)pc.m_tile * 128 + rbase + 16 * j;
        const TX* row = x + m * ld;
        if (m < M && vec && k + 4 <= K) {
          dst[j] = __ldg(reinterpret_cast<const Raw*>(row + k));
        } else {
          TX v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (m < M && k + e < K) ? row[k + e] : from_float<TX>(0.f);
          dst[j] = *reinterpret_cast<const Raw*>(v);
        }
      }
    };
    auto emit_step = [&](const Raw (&src)[8], int it) {
      uint8_t* dst = pc.stage(it) + off;
      const int ks = pc.ks0 + it;
      uint8_t* blk = emit ? bf16_pack + packed_block_index(pc.m_tile, ks, pack_row_blocks) * kBlockBytes + off : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v[4];
        unpack(src[j], v);
        const int ro = j * 16 * 128;
        if constexpr (kLo) {
          uint2 hi, lo;
          split_f16x4(v, hi, lo);
          *reinterpret_cast<uint2*>(dst + ro) = hi;
          *reinterpret_cast<uint2*>(dst + kBlockBytes + ro) = lo;
        } else {
          // exact for bf16 inputs inside the f16 range (clamped like the split path); no residual part
          const __half2 h01 = __floats2half2_rn(fminf(fmaxf(v[0], -65504.f), 65504.f), fminf(fmaxf(v[1], -65504.f), 65504.f));
          const __half2 h23 = __floats2half2_rn(fminf(fmaxf(v[2], -65504.f), 65504.f), fminf(fmaxf(v[3], -65504.f), 65504.f));
          *reinterpret_cast<uint2*>(dst + ro) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        }
        if (emit) *reinterpret_cast<uint2*>(blk + ro) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
      }
    };
    Raw qa[8], qb[8];
    load(qa, pc.ks0);
    for (int it = 0; it < pc.n_it; it += 2) {
      if (it + 1 < pc.n_it) load(qb, pc.ks0 + it + 1);
      pc.mark(9, it);
      pc.wait_empty(it);
      pc.mark(1, it);
      emit_step(qa, it);
      pc.arrive_full(it);
      pc.mark(2, it);
      if (it + 2 < pc.n_it) load(qa, pc.ks0 + it + 2);
      if (it + 1 < pc.n_it) {
        pc.wait_empty(it + 1);
        emit_step(qb, it + 1);
        pc.arrive_full(it + 1);
      }
    }
  }
};

using RowSplitProducerF16 = RowSplitProducerF16T<float>;

// C[m * ldc + n] = bf16(acc): plain row-major bf16 result (the gradient of bf16 activations), transposed through shared
// memory like StoreRowMajorEpi so that a warp store covers whole sectors
struct StoreRowMajorBf16Epi {
  __nv_bfloat16* C;
  int64_t ldc;
  int M, N;
  static constexpr int kScratchBytes = kTransposeScratchBytes;
  struct State {};
  __device__ void begin(State&, const EpiCtx&) const {}
  __device__ void end(State&, const EpiCtx&) const {}
  __device__ void chunk(State&, const EpiCtx& ctx, int n, const float (&acc)[32]) const {
    const int m0 = ctx.m - (ctx.t & 31);
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 7) == 0);
    warp_transposed_chunk(ctx, acc, N - n, [&](int r, int c, float4 v) {
      const int m = m0 + r, col = n + c;
      if (m >= M || col >= N) return;
      __nv_bfloat16* dst = C + (int64_t)m * ldc + col;
      if (vec_ok && col + 4 <= N) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) dst[j] = __float2bfloat16(e[j]);
      }
    });
  }
};

}  // namespace tc
}  // namespace s2t
