// sm_100a primitives for the tensor-core kernels: mbarrier, bulk async copies (TMA engine,
// SASS UBLKCP), tcgen05 MMA with TMEM accumulators, and the descriptor / tile-layout helpers.
//
// Operand tiles ("blocks") in shared memory are 128 rows x 64 bf16 (128 B per row, 16 KB per
// block), K-major with the 128-byte swizzle the UMMA descriptors expect: the 16-byte chunk c of
// row r sits at chunk position (c ^ (r & 7)).  Blocks are stored in global memory already in
// that image (pack kernels in tc_pack.cu), so one contiguous cp.async.bulk brings a block --
// or several adjacent blocks -- in; no tensor map is needed.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace s2t {
namespace tc {

constexpr int kBlockRows = 128;                                   // rows of one operand block
constexpr int kBlockK = 64;                                       // bf16 elements along K per block
constexpr int kBlockBytes = kBlockRows * kBlockK * 2;             // 16384
constexpr int kUmmaK = 16;                                        // K per tcgen05.mma (bf16)

// byte offset of element (r, k) inside one block, r < 128, k < 64
__host__ __device__ __forceinline__ uint32_t block_elem_offset(int r, int k) {
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1));
}
// byte offset of the 16-byte chunk c (8 elements) of row r
__host__ __device__ __forceinline__ uint32_t block_chunk_offset(int r, int c) {
  return (uint32_t)(r * 128 + (((c ^ (r & 7)) & 7) << 4));
}

// A "packed operand" X (rows x K) is an array of blocks ordered k-block major, row-block minor:
//   block(rb, kb) at index kb * row_blocks + rb.
__host__ __device__ __forceinline__ size_t packed_block_index(int rb, int kb, int row_blocks) {
  return (size_t)kb * row_blocks + rb;
}
__host__ __device__ __forceinline__ size_t packed_bytes(int64_t rows, int64_t K) {
  int64_t rbk = (rows + kBlockRows - 1) / kBlockRows, kbk = (K + kBlockK - 1) / kBlockK;
  return (size_t)rbk * kbk * kBlockBytes;
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Spin on the phase parity.  A wait that does not complete within ~2^28 probes means a
// protocol bug; trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spins > (1u << 28)) __trap();
  }
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream is still draining.  Everything before griddep_wait() (shared-memory
// carve-up, barrier init, TMEM allocation) overlaps the predecessor's tail; griddep_wait() returns once the predecessor
// grid has completed and its writes are visible.  griddep_launch_dependents() lets the NEXT kernel's CTAs be scheduled as
// soon as every CTA of this grid has issued it (called after the own wait, so "my successor runs" implies "my
// predecessor is complete").  Both are no-ops in a kernel that was launched without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Non-blocking probe of a phase parity (a role that serves two pipelines polls both instead of blocking on one).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// Wait on a barrier whose arrivals come from the peer CTA of a pair (same default semantics as the remote arrive
// below: a cluster-scope release / acquire pair costs the forwarding thread most of a microsecond per stage).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spins > (1u << 28)) __trap();
  }
}
// arrive on the mbarrier at this shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// ---- bulk async copy global -> shared (TMA engine) --------------------------------------------
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D tiled tensor copy global -> shared through a CUtensorMap (cuTensorMapEncodeTiled; the map lives in kernel
// parameter space, __grid_constant__): one instruction moves a box of rows x columns, c0 = first column (innermost
// dimension), c1 = first row; elements outside the tensor arrive as zeros and still count in the transaction bytes.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// same copy delivered to the same shared-memory offset (and mbarrier) of every CTA in cta_mask of the cluster
__device__ __forceinline__ void bulk_copy_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                        uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of 2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// CTA-pair variants: one warp of EACH CTA of the pair executes them; both tensor memories get the same columns
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of 2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors -------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset (unused: one swizzle atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset between 8-row groups [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: kind::f16, A = B = bf16, D = fp32, both K-major, M x N tile
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format: f32
  d |= 1u << 7;                      // A format: bf16
  d |= 1u << 10;                     // B format: bf16
  d |= (uint32_t)(N >> 3) << 17;     // N
  d |= (uint32_t)(M >> 4) << 24;     // M
  return d;
}
// same instruction kind with IEEE half operands (11 significant bits): used for the hi/lo split that carries
// fp32-level accuracy through three MMAs at the full 16-bit rate
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format: f32; A / B format fields 0 = f16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread (kind::f16 covers f16 and bf16 operands)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// tf32 variant: operands are fp32 in smem (same 128-byte swizzled rows, 32 elements per row), K = 8 per MMA
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format: f32
  d |= 2u << 7;                      // A format: tf32
  d |= 2u << 10;                     // B format: tf32
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// CTA-pair MMA (cta_group::2), issued by one thread of the leader CTA: M = 256 (128 rows of A from each CTA's
// shared memory at the same offset), B = N/2 rows from each CTA, D = 128 lanes x N columns in each CTA's TMEM.
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// completion of the pair's MMAs: one arrival on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// fp32 -> (hi, lo) halves with hi + lo = x to ~22 significant bits (|x| clamped to the f16 range)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// four floats -> four hi halves and four lo halves (8 bytes each).  Packed conversions only (cvt.rn.f16x2.f32 and the
// half2 -> float2 widening): the scalar cvt.f16.f32 issues on the narrow conversion pipe and made the producers that
// split a whole operand stage per k-step the slowest role of their kernels (role timeline, profiles/README.md).
__device__ __forceinline__ void split_f16x4(const float (&x)[4], uint2& hi, uint2& lo) {
  float c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = fminf(fmaxf(x[j], -65504.f), 65504.f);
  const __half2 h01 = __floats2half2_rn(c[0], c[1]), h23 = __floats2half2_rn(c[2], c[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(c[0] - f01.x, c[1] - f01.y), l23 = __floats2half2_rn(c[2] - f23.x, c[3] - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// round-to-nearest (ties away from zero) fp32 -> tf32; the MMA itself would truncate the low 13 mantissa bits,
// which is biased.  Same result as cvt.rna.tf32.f32 for finite values, but two full-rate integer instructions
// (the cvt issues on a narrow pipe: "math pipe throttle" was the producers' top stall reason in ncu).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#endif  // __CUDACC__

}  // namespace tc
}  // namespace s2t
