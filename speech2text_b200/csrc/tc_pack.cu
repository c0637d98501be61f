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

// Row-major fp32 (rows, K) -> bf16 packed operand, one 128 x 64 block per CTA, plus the column sums of the
// source (the bias gradient of a linear layer falls out of packing dy).  16 threads cover the 64 columns
// of a row (one float4 each), 16 rows per pass; the block's column sums meet in shared memory and leave
// as 64 atomics.
__global__ void __launch_bounds__(256) pack_rows_colsum_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ src2, int64_t ld, int rows, int K,
                                                               int row_blocks, uint8_t* __restrict__ dst,
                                                               float* __restrict__ col_sum) {
  __shared__ float part[16][65];
  const int rb = blockIdx.x, kb = blockIdx.y;
  const int c = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int k = kb * 64 + c * 4;
  const bool vec = ((ld & 3) == 0) && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(src2)) & 15) == 0);
  uint8_t* blk = dst + packed_block_index(rb, kb, row_blocks) * kBlockBytes;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // blockIdx.z = upper / lower 64 rows of the block: 128 x 64 tiles gave 1600 CTAs for the encoder-side dy of c3,
  // 1.35 waves of the 1184 resident ones, i.e. two rounds for 1.35 rounds of work
#pragma unroll
  for (int pp = 0; pp < 4; ++pp) {
    const int pass = (int)blockIdx.z * 4 + pp;
    const int rr = pass * 16 + rl;
    const int64_t r = (int64_t)rb * 128 + rr;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
      const float* p = src + r * ld + k;
      if (vec && k + 4 <= K) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        if (src2) {  // the operand is the sum of two gradient terms that were never added in memory
          const float4 u = __ldg(reinterpret_cast<const float4*>(src2 + r * ld + k));
          v[0] += u.x; v[1] += u.y; v[2] += u.z; v[3] += u.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (k + e < K) ? __ldg(p + e) + (src2 ? __ldg(src2 + r * ld + k + e) : 0.f) : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += v[e];
    *reinterpret_cast<uint2*>(blk + block_chunk_offset(rr, c >> 1) + (c & 1) * 8) =
        make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
  if (col_sum == nullptr) return;
#pragma unroll
  for (int e = 0; e < 4; ++e) part[rl][c * 4 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += part[i][threadIdx.x];
    const int col = kb * 64 + threadIdx.x;
    if (col < K && t != 0.f) atomicAdd(col_sum + col, t);
  }
}

// several small matrices (the weights of one module) in ONE launch: one thread per 16-byte chunk of any job
__global__ void pack_jobs_kernel(PackJobs jobs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= jobs.first[jobs.n]) {
    const int64_t f = i - jobs.first[jobs.n];
    if (f < jobs.fill_n) jobs.fill[f] = jobs.fill_value;
    return;
  }
  int j = 0;
#pragma unroll
  for (int q = 1; q < PackJobs::kMax; ++q)
    if (q < jobs.n && i >= jobs.first[q]) j = q;
  const PackJob& job = jobs.job[j];
  const int64_t li = i - jobs.first[j];
  const int64_t rows_pad = (int64_t)job.row_blocks * 128, chunks = (int64_t)job.k_blocks * 8;
  int64_t r, ck;
  if (job.col_stride == 1) {  // consecutive threads along K
    ck = li % chunks;
    r = li / chunks;
  } else {  // transposed source: consecutive threads along the rows
    r = li % rows_pad;
    ck = li / rows_pad;
  }
  const int per = (job.kind == 0 || job.kind >= 3) ? 8 : 4;  // elements per 16-byte chunk
  const int k0 = (int)ck * per;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = (e < per && r < job.rows && k < job.K) ? __ldg(job.src + r * job.row_stride + (int64_t)k * job.col_stride) : 0.f;
  }
  const int rb = (int)(r >> 7), kb = (int)(ck >> 3);
  uint8_t* out = job.dst + packed_block_index(rb, kb, job.row_blocks) * kBlockBytes + block_chunk_offset((int)(r & 127), (int)(ck & 7));
  if (job.kind == 0) {
    *reinterpret_cast<uint4*>(out) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  } else if (job.kind >= 3) {
    const float a[4] = {v[0] * job.scale, v[1] * job.scale, v[2] * job.scale, v[3] * job.scale};
    const float b[4] = {v[4] * job.scale, v[5] * job.scale, v[6] * job.scale, v[7] * job.scale};
    uint2 h0, l0, h1, l1;
    split_f16x4(a, h0, l0);
    split_f16x4(b, h1, l1);
    *reinterpret_cast<uint4*>(out) = job.kind == 3 ? make_uint4(h0.x, h0.y, h1.x, h1.y) : make_uint4(l0.x, l0.y, l1.x, l1.y);
  } else {
    float4 o;
    float* po = &o.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float big = round_tf32(v[e]);
      po[e] = job.kind == 1 ? big : round_tf32(v[e] - big);
    }
    *reinterpret_cast<float4*>(out) = o;
  }
}

// batched, optionally exponentiated packing: one thread per 16-byte chunk
template <bool kF32>
__global__ void pack_batched_kernel(PackSpec p, uint8_t* __restrict__ dst, uint8_t* __restrict__ dst_small,
                                    uint8_t* __restrict__ dst_bf16, int bf16_k_blocks) {
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
    if (dst_bf16) {  // the same values as a bf16 operand (64 per block row): this thread owns half a 16-byte chunk
      const size_t o16 = packed_block_index((int)rb, ck >> 4, (int)(((int64_t)p.batches * p.rows_pad) >> 7)) * kBlockBytes +
                         block_chunk_offset((int)(rp & 127), (ck >> 1) & 7) + (ck & 1) * 8;
      if ((ck >> 4) < bf16_k_blocks)
        *reinterpret_cast<uint2*>(dst_bf16 + o16) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  } else {
    uint4 out;
    out.x = pack_bf16x2(v[0], v[1]);
    out.y = pack_bf16x2(v[2], v[3]);
    out.z = pack_bf16x2(v[4], v[5]);
    out.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + off) = out;
  }
}

// batched hi / lo f16 split of scale * f(x) with the optional exp, plus bf16 f(x): one thread per 16-byte chunk
__global__ void pack_batched_f16split_kernel(PackSpec p, float scale, uint8_t* __restrict__ dst_hi,
                                             uint8_t* __restrict__ dst_lo, uint8_t* __restrict__ dst_bf16) {
  const int chunks_per_row = p.k_blocks * 8;
  const int64_t total = (int64_t)p.batches * p.rows_pad * chunks_per_row;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ck = (int)(i % chunks_per_row);
  const int64_t rp = i / chunks_per_row;
  const int batch = (int)(rp / p.rows_pad), r = (int)(rp % p.rows_pad);
  const int k0 = ck * 8;
  const bool row_ok = r < p.rows;
  const float* src = p.src + (int64_t)batch * p.batch_stride + (int64_t)r * p.row_stride;
  const float sub = (row_ok && p.row_sub) ? __ldg(p.row_sub + (int64_t)batch * p.rows + r) : 0.f;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = 0.f;
    if (row_ok && k0 + j < p.K) {
      x = __ldg(src + k0 + j);
      if (p.row_sub) x = expf(x - sub);
    }
    v[j] = x;
  }
  const size_t off = packed_block_index((int)(rp >> 7), ck >> 3, (int)(((int64_t)p.batches * p.rows_pad) >> 7)) * kBlockBytes +
                     block_chunk_offset((int)(rp & 127), ck & 7);
  const float a[4] = {v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale};
  const float b[4] = {v[4] * scale, v[5] * scale, v[6] * scale, v[7] * scale};
  uint2 h0, l0, h1, l1;
  split_f16x4(a, h0, l0);
  split_f16x4(b, h1, l1);
  *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
  if (dst_bf16)
    *reinterpret_cast<uint4*>(dst_bf16 + off) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

}  // namespace

int pack_f16_split(const PackSpec& p, float scale, uint8_t* dst_hi, uint8_t* dst_lo, uint8_t* dst_bf16,
                   cudaStream_t stream) {
  S2T_REQUIRE(p.rows_pad % 128 == 0 && p.rows_pad >= p.rows && p.k_blocks * 64 >= p.K, "pack_f16_split: bad padding");
  const int64_t total = (int64_t)p.batches * p.rows_pad * p.k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_batched_f16split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, scale, dst_hi, dst_lo, dst_bf16);
  return check_launch("pack_batched_f16split_kernel");
}

int pack_bf16(const PackSpec& p, uint8_t* dst, cudaStream_t stream) {
  S2T_REQUIRE(p.rows_pad % 128 == 0 && p.rows_pad >= p.rows && p.k_blocks * 64 >= p.K, "pack_bf16: bad padding");
  const int64_t total = (int64_t)p.batches * p.rows_pad * p.k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_batched_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, dst, nullptr, nullptr, 0);
  return check_launch("pack_batched_kernel<bf16>");
}

int pack_rows_colsum(const float* src, const float* src2, int64_t ld, int rows, int K, int row_blocks, int k_blocks,
                     uint8_t* dst, float* col_sum, cudaStream_t stream) {
  S2T_REQUIRE(row_blocks * 128 >= rows && k_blocks * 64 >= K, "pack_rows_colsum: padded dims too small");
  if (row_blocks == 0 || k_blocks == 0) return 0;
  S2T_REQUIRE(k_blocks <= 65535, "pack_rows_colsum: K too large");
  ProfScope prof("pack_operand_kernel", stream);
  pack_rows_colsum_kernel<<<dim3((unsigned)row_blocks, (unsigned)k_blocks, 2), 256, 0, stream>>>(src, src2, ld, rows, K, row_blocks, dst,
                                                                                            col_sum);
  return check_launch("pack_rows_colsum_kernel");
}

int pack_jobs(const PackJob* list, int n, cudaStream_t stream, float* fill, int64_t fill_n, float fill_value) {
  S2T_REQUIRE(n >= 1 && n <= PackJobs::kMax, "pack_jobs: %d jobs", n);
  PackJobs jobs;
  jobs.n = n;
  jobs.first[0] = 0;
  for (int j = 0; j < n; ++j) {
    const PackJob& q = list[j];
    S2T_REQUIRE(q.row_blocks * 128 >= q.rows && q.k_blocks * ((q.kind == 0 || q.kind >= 3) ? 64 : 32) >= q.K,
                "pack_jobs: padded dims too small");
    jobs.job[j] = q;
    jobs.first[j + 1] = jobs.first[j] + (int64_t)q.row_blocks * 128 * q.k_blocks * 8;
  }
  jobs.fill = fill;
  jobs.fill_n = fill ? fill_n : 0;
  jobs.fill_value = fill_value;
  const int64_t total = jobs.first[n] + jobs.fill_n;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_jobs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(jobs);
  return check_launch("pack_jobs_kernel");
}

int pack_f32_split(const PackSpec& p, uint8_t* dst_big, uint8_t* dst_small, cudaStream_t stream, uint8_t* dst_bf16,
                   int bf16_k_blocks) {
  S2T_REQUIRE(p.rows_pad % 128 == 0 && p.rows_pad >= p.rows && p.k_blocks * 32 >= p.K, "pack_f32_split: bad padding");
  const int64_t total = (int64_t)p.batches * p.rows_pad * p.k_blocks * 8;
  if (total == 0) return 0;
  ProfScope prof("pack_operand_kernel", stream);
  pack_batched_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, dst_big, dst_small, dst_bf16,
                                                                                 bf16_k_blocks);
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
  const bool pair = getenv("S2T_GEMM_CLUSTER") && atoi(getenv("S2T_GEMM_CLUSTER")) == 2;  // 2-CTA clusters, B multicast
  if (bn == 128) {
    if (pair)
      return tc::launch_gemm_stream<128, 4, false, 0, 2>(a, pb, b_row_blocks, m_tiles, (N + 127) / 128, k_blocks, k_splits, epi,
                                                         st, "tc_gemm_debug_128_pair");
    return tc::launch_gemm_stream<128, 4, false, 0>(a, pb, b_row_blocks, m_tiles, (N + 127) / 128, k_blocks, k_splits, epi, st,
                                          "tc_gemm_debug_128");
  }
  if (pair)
    return tc::launch_gemm_stream<256, 4, false, 0, 2>(a, pb, b_row_blocks, m_tiles, (N + 255) / 256, k_blocks, k_splits, epi, st,
                                                       "tc_gemm_debug_256_pair");
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
