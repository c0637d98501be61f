// fp32 SIMT contraction template with fused operand producers and epilogues.
//
//     C(batch, m, n) = sum_k A(batch, m, k) * B(batch, n, k)
//
// The strict-fp32 mode of the hot path (north_star parity: 1e-5 loss / 1e-4
// grads against the reference's fp32 CPU result) cannot go through bf16 tensor
// cores, so every contraction of that mode -- the normaliser bmm of the simple
// loss, its backward (A.7) and the joiner's V->I->V projections
// (/root/reference/model/joiner/joiner.py:176-178) -- is an instance of this
// kernel.  Operands are never materialised: A and B are *functors* that build
// the element on the fly (exp(x - rowmax), act(am + lm[ranges]), ...), and the
// epilogue functor consumes the accumulator (log + gather, bias, atomics ...).
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread as two
// 4-wide groups 64 apart (conflict-free float4 LDS), global->register->shared
// double buffering with one barrier per k-tile.
//
// Functor contract:
//   struct Operand {
//     struct Row { ... };                                   // per-row state, cached in smem
//     __device__ Row row(int batch, int idx) const;         // idx = m (A) or n (B), in bounds
//     __device__ float at(const Row&, int k) const;         // k in bounds
//   };
//   struct Epilogue { __device__ void operator()(int batch, int m, int n, float acc) const; };
// K_CONTIG says which index is contiguous in memory for that operand and only
// selects the thread->element mapping of the tile load (coalescing).
#pragma once
#include "common.cuh"

namespace s2t {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmThreads = 256;
constexpr int kGemmLd = kGemmBM + 4;

template <bool A_KCONTIG, bool B_KCONTIG, class AF, class BF, class EF>
__global__ void __launch_bounds__(kGemmThreads)
sgemm_kernel(int M, int N, int K, int k_splits, AF af, BF bf, EF ef) {
  __shared__ __align__(16) float As[2][kGemmBK][kGemmLd];
  __shared__ __align__(16) float Bs[2][kGemmBK][kGemmLd];
  __shared__ typename AF::Row a_rows[kGemmBM];
  __shared__ typename BF::Row b_rows[kGemmBN];

  const int tid = threadIdx.x;
  const int batch = blockIdx.z / k_splits;
  const int split = blockIdx.z % k_splits;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;

  // K range of this split, in whole k-tiles
  const int k_tiles = (K + kGemmBK - 1) / kGemmBK;
  const int tiles_per = (k_tiles + k_splits - 1) / k_splits;
  const int kt_begin = split * tiles_per;
  const int kt_end = min(k_tiles, kt_begin + tiles_per);

  for (int i = tid; i < kGemmBM; i += kGemmThreads) {
    if (m0 + i < M) a_rows[i] = af.row(batch, m0 + i);
  }
  for (int i = tid; i < kGemmBN; i += kGemmThreads) {
    if (n0 + i < N) b_rows[i] = bf.row(batch, n0 + i);
  }
  __syncthreads();

  float ra[8], rb[8];
  auto fetch = [&](int kt) {
    const int kbase = kt * kGemmBK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int mm, kk;
      if (A_KCONTIG) {
        kk = tid % kGemmBK;
        mm = tid / kGemmBK + (kGemmThreads / kGemmBK) * j;
      } else {
        mm = tid % kGemmBM;
        kk = tid / kGemmBM + (kGemmThreads / kGemmBM) * j;
      }
      ra[j] = (m0 + mm < M && kbase + kk < K) ? af.at(a_rows[mm], kbase + kk) : 0.f;
      int nn, kb;
      if (B_KCONTIG) {
        kb = tid % kGemmBK;
        nn = tid / kGemmBK + (kGemmThreads / kGemmBK) * j;
      } else {
        nn = tid % kGemmBN;
        kb = tid / kGemmBN + (kGemmThreads / kGemmBN) * j;
      }
      rb[j] = (n0 + nn < N && kbase + kb < K) ? bf.at(b_rows[nn], kbase + kb) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int mm, kk;
      if (A_KCONTIG) {
        kk = tid % kGemmBK;
        mm = tid / kGemmBK + (kGemmThreads / kGemmBK) * j;
      } else {
        mm = tid % kGemmBM;
        kk = tid / kGemmBM + (kGemmThreads / kGemmBM) * j;
      }
      As[buf][kk][mm] = ra[j];
      int nn, kb;
      if (B_KCONTIG) {
        kb = tid % kGemmBK;
        nn = tid / kGemmBK + (kGemmThreads / kGemmBK) * j;
      } else {
        nn = tid % kGemmBN;
        kb = tid / kGemmBN + (kGemmThreads / kGemmBN) * j;
      }
      Bs[buf][kb][nn] = rb[j];
    }
  };

  const int tx = tid % 16, ty = tid / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (kt_begin < kt_end) {
    fetch(kt_begin);
    stash(0);
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
      const int buf = (kt - kt_begin) & 1;
      if (kt + 1 < kt_end) fetch(kt + 1);
#pragma unroll
      for (int kk = 0; kk < kGemmBK; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kt + 1 < kt_end) {
        stash(buf ^ 1);
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) ef(batch, m, n, acc[i][j]);
    }
  }
}

template <bool A_KCONTIG, bool B_KCONTIG, class AF, class BF, class EF>
int launch_sgemm(int batches, int M, int N, int K, int k_splits, const AF& af, const BF& bf,
                 const EF& ef, cudaStream_t stream, const char* what) {
  if (batches <= 0 || M <= 0 || N <= 0) return 0;
  if (k_splits < 1) k_splits = 1;
  dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM, batches * k_splits);
  S2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large (%u, %u)", what, grid.y, grid.z);
  ProfScope prof(what, stream);
  sgemm_kernel<A_KCONTIG, B_KCONTIG, AF, BF, EF><<<grid, kGemmThreads, 0, stream>>>(M, N, K, k_splits, af, bf, ef);
  return check_launch(what);
}

// ---- plain strided operands ------------------------------------------------
// element(batch, idx, k) = p[batch*bs + idx*is + k*ks]
struct StridedOperand {
  const float* p;
  int64_t bs, is, ks;
  struct Row { const float* p; };
  __device__ Row row(int batch, int idx) const { return Row{p + batch * bs + idx * is}; }
  __device__ float at(const Row& r, int k) const { return __ldg(r.p + k * ks); }
};

}  // namespace s2t
