// fp32 row/column-strided matrix -> bf16 "packed operand" (tc_prims.cuh): 128 x 64 blocks in the
// swizzled shared-memory image, k-block major, zero padded.  Used for the weights (re-packed every
// step: they change under training) and for fp32 activations that feed a tensor-core contraction.
#include "tc_gemm.cuh"

namespace s2t {
namespace tc {
namespace {

// one thread per 16-byte chunk (8 consecutive k of one row)
template <bool kRowsFastest>
__global__ void pack_operand_kernel(const float* __restrict__ src, int64_t row_stride, int64_t col_stride,
                                    int rows, int K, int row_blocks, int k_blocks, uint8_t* __restrict__ dst) {
  const int64_t total = (int64_t)row_blocks * 128 * k_blocks * 8;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t r, ck;  // row, chunk index along K
  if (kRowsFastest) {
    r = i % ((int64_t)row_blocks * 128);
    ck = i / ((int64_t)row_blocks * 128);
  } else {
    ck = i % ((int64_t)k_blocks * 8);
    r = i / ((int64_t)k_blocks * 8);
  }
  const int k0 = (int)ck * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    v[j] = (r < rows && k < K) ? __ldg(src + r * row_stride + (int64_t)k * col_stride) : 0.f;
  }
  uint4 out;
  out.x = pack_bf16x2(v[0], v[1]);
  out.y = pack_bf16x2(v[2], v[3]);
  out.z = pack_bf16x2(v[4], v[5]);
  out.w = pack_bf16x2(v[6], v[7]);
  const int rb = (int)(r / 128), kb = k0 / 64;
  uint8_t* blk = dst + packed_block_index(rb, kb, row_blocks) * kBlockBytes;
  *reinterpret_cast<uint4*>(blk + block_chunk_offset((int)(r % 128), (k0 % 64) / 8)) = out;
}

// fp32 elements: one thread per 16-byte chunk (4 consecutive k of one row), 32 k per block row
__global__ void pack_operand_f32_kernel(const float* __restrict__ src, int64_t row_stride, int rows, int K,
                                        int row_blocks, int k_blocks, int part, uint8_t* __restrict__ dst) {
  const int64_t total = (int64_t)row_blocks * 128 * k_blocks * 8;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t ck = i % ((int64_t)k_blocks * 8);
  const int64_t r = i / ((int64_t)k_blocks * 8);
  const int k0 = (int)ck * 4;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = (r < rows && k0 + j < K) ? __ldg(src + r * row_stride + k0 + j) : 0.f;
    float big = round_tf32(x);
    v[j] = part == 0 ? big : round_tf32(x - big);
  }
  const int rb = (int)(r / 128), kb = k0 / 32;
  uint8_t* blk = dst + packed_block_index(rb, kb, row_blocks) * kBlockBytes;
  *reinterpret_cast<float4*>(blk + block_chunk_offset((int)(r % 128), (k0 % 32) / 4)) = make_float4(v[0], v[1], v[2], v[3]);
}

// batched, optionally exponentiated packing: one thread per 16-byte chunk
template <bool kF32>
__global__ void pack_batched_kernel(PackSpec p, uint8_t* __restrict__ dst, uint8_t* __restrict__ dst_small) {
  constexpr int kPer = kF32 ? 4 : 8;  // elements per chunk
  const int chunks_per_row = p.k_blocks * 8;
  const int64_t total = (int64_t)p.batches * p.rows_pad * chunks_per_row;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ck = (int)(i % chunks_per_row);
  const int64_t rp = i / chunks_per_row;  // padded global row
  const int batch = (int)(rp / p.rows_pad), r = (int)(rp % p.rows_pad);
  const int k0 = ck * kPer;
  float v[kPer];
  const bool row_ok = r < p.rows;
  const float* src = p.src + (int64_t)batch * p.batch_stride + (int64_t)r * p.row_stride;
  const float sub = (row_ok && p.row_sub) ? __ldg(p.row_sub + (int64_t)batch * p.rows + r) : 0.f;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    float x = 0.f;
    if (row_ok && k0 + j < p.K) {
      x = __ldg(src + k0 + j);
      if (p.row_sub) x = expf(x - sub);
    }
    v[j] = x;
  }
  const int64_t rb = rp >> 7;
  const int kb = ck >> 3;
  const size_t off = packed_block_index((int)rb, kb, (int)(((int64_t)p.batches * p.rows_pad) >> 7)) * kBlockBytes +
                     block_chunk_offset((int)(rp & 127), ck & 7);
  if constexpr (kF32) {
    float4 big, small;
    big.x = round_tf32(v[0]); big.y = round_tf32(v[1]); big.z = round_tf32(v[2]); big.w = round_tf32(v[3]);
    small.x = round_tf32(v[0] - big.x); small.y = round_tf32(v[1] - big.y);
    small.z = round_tf32(v[2] - big.z); small.w = round_tf32(v[3] - big.w);
    *reinterpret_cast<float4*>(dst + off) = big;
    *reinterpret_cast<float4*>(dst_small + off) = small;
  } else {
    uint4 out;
    out.x = pack_bf16x2(v[0], v[1]);
    out.y = pack_bf16x2(v[2], v[3]);
    out.z = pack_bf16x2(v[4], v[5]);
    out.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + off) = out;
  }
}

}  // namespace

int pack_bf16(const PackSpec& p, uint8_t* dst, cudaStream_t stream) {
  S2T_REQUIRE(p.rows_pad % 128 == 0 && p.rows_pad >= p.rows && p.k_blocks * 64 >= p.K, "pack_bf16: bad padding");
  const int64_t total = (int64_t)p.batches * p.rows_pad * p.k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_batched_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, dst, nullptr);
  return check_launch("pack_batched_kernel<bf16>");
}

int pack_f32_split(const PackSpec& p, uint8_t* dst_big, uint8_t* dst_small, cudaStream_t stream) {
  S2T_REQUIRE(p.rows_pad % 128 == 0 && p.rows_pad >= p.rows && p.k_blocks * 32 >= p.K, "pack_f32_split: bad padding");
  const int64_t total = (int64_t)p.batches * p.rows_pad * p.k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_batched_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, dst_big, dst_small);
  return check_launch("pack_batched_kernel<f32>");
}

int pack_operand_f32(const float* src, int64_t row_stride, int rows, int K, int row_blocks, int k_blocks,
                     int part, uint8_t* dst, cudaStream_t stream) {
  S2T_REQUIRE(row_blocks * 128 >= rows && k_blocks * 32 >= K, "pack_operand_f32: padded dims too small");
  const int64_t total = (int64_t)row_blocks * 128 * k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_operand_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, row_stride, rows, K, row_blocks,
                                                                               k_blocks, part, dst);
  return check_launch("pack_operand_f32_kernel");
}

int pack_operand(const float* src, int64_t row_stride, int64_t col_stride, int rows, int K, int row_blocks,
                 int k_blocks, uint8_t* dst, cudaStream_t stream) {
  S2T_REQUIRE(row_blocks * 128 >= rows && k_blocks * 64 >= K, "pack_operand: padded dims too small");
  const int64_t total = (int64_t)row_blocks * 128 * k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (col_stride == 1) {
    pack_operand_kernel<false><<<grid, 256, 0, stream>>>(src, row_stride, col_stride, rows, K, row_blocks, k_blocks, dst);
  } else {
    pack_operand_kernel<true><<<grid, 256, 0, stream>>>(src, row_stride, col_stride, rows, K, row_blocks, k_blocks, dst);
  }
  return check_launch("pack_operand_kernel");
}

}  // namespace tc
}  // namespace s2t

using namespace s2t;

// Debug / test entry: C (M x N, fp32) = A (M x K) * B (N x K)^T with bf16-rounded operands on the
// tensor cores.  ws must hold packed A + packed B.  bn in {128, 256}.
extern "C" size_t s2t_tc_gemm_workspace_bytes(int M, int N, int K) {
  return tc::packed_bytes(M, K) + tc::packed_bytes(((N + 255) / 256) * 256, K) + 256;
}

extern "C" int s2t_tc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int bn, int k_splits,
                           void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  S2T_REQUIRE(bn == 128 || bn == 256, "tc_gemm: bn must be 128 or 256");
  uint8_t* pa = (uint8_t*)ws;
  uint8_t* pb = pa + tc::packed_bytes(M, K);
  const int n_pad = ((N + 255) / 256) * 256;
  const int b_row_blocks = n_pad / 128;
  const int m_tiles = (M + 127) / 128, k_blocks = (K + 63) / 64;
  if (int rc = tc::pack_operand(A, K, 1, M, K, m_tiles, k_blocks, pa, st)) return rc;
  if (int rc = tc::pack_operand(B, K, 1, N, K, b_row_blocks, k_blocks, pb, st)) return rc;
  tc::BulkA a{pa, m_tiles};
  if (k_splits > 1) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
  tc::StoreRowMajorEpi epi{C, N, M, N, k_splits > 1};
  if (bn == 128) {
    return tc::launch_gemm_stream<128, 4, false, 0>(a, pb, b_row_blocks, m_tiles, (N + 127) / 128, k_blocks, k_splits, epi, st,
                                          "tc_gemm_debug_128");
  }
  return tc::launch_gemm_stream<256, 4, false, 0>(a, pb, b_row_blocks, m_tiles, (N + 255) / 256, k_blocks, k_splits, epi, st,
                                        "tc_gemm_debug_256");
}

// Debug / test entry for MN-major operands: C (M x N) = At^T Bt with At (K x M), Bt (K x N) fp32 row-major.
extern "C" size_t s2t_tc_gemm_mn_workspace_bytes(int M, int N, int K) {
  const int rb = (K + 127) / 128;
  return (size_t)rb * (((M + 127) / 128) * 2 + ((N + 255) / 256) * 4) * tc::kBlockBytes + 256;
}

extern "C" int s2t_tc_gemm_mn(const float* At, const float* Bt, float* C, int M, int N, int K, int k_splits,
                              unsigned lbo, unsigned sbo, unsigned kadv, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int rb = (K + 127) / 128;
  const int a_groups = ((M + 127) / 128) * 2, b_groups = ((N + 255) / 256) * 4;
  uint8_t* pa = (uint8_t*)ws;
  uint8_t* pb = pa + (size_t)rb * a_groups * tc::kBlockBytes;
  if (int rc = tc::pack_operand(At, M, 1, K, M, rb, a_groups, pa, st)) return rc;
  if (int rc = tc::pack_operand(Bt, N, 1, K, N, rb, b_groups, pb, st)) return rc;
  tc::BulkA a{pa, rb};
  if (k_splits > 1) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
  tc::StoreRowMajorEpi epi{C, N, M, N, k_splits > 1};
  tc::MnDebug mn;
  if (lbo) mn.lbo_bytes = lbo;
  if (sbo) mn.sbo_bytes = sbo;
  if (kadv) mn.k_advance_bytes = kadv;
  return tc::launch_gemm_stream<256, 4, true, 0>(a, pb, rb, (M + 127) / 128, (N + 255) / 256, (K + 63) / 64, k_splits, epi,
                                              st, "tc_gemm_mn_debug", mn);
}
