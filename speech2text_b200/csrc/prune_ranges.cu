// Prune-range selection: for every frame pick the window of `s_range` symbol
// positions that carries the most occupation mass of the simple lattice, then
// make the window starts monotone and gap-free.
//
// Replaces k2.get_rnnt_prune_ranges as called from
// /root/reference/model/joiner/joiner.py:112-117  (k2 semantics: SURVEY.md A.4).
// Integer output, bit-exact contract: identical (px_grad, py_grad) inputs, fp32
// left-to-right summation over the window, first-maximum tie-break.
//
// Two launches:
//   prune_argmax_kernel  thread (frame t, candidate segment): argmax over s of the window score.  Reads
//            are coalesced along t (the occupation arrays are t-contiguous); the candidates of a frame
//            are split over 8 segments so that B*T*8 threads share the latency-bound scan, and merged in
//            segment order with a strict '>' (first maximum wins, as in a serial scan).  The raw window
//            starts are parked in ranges[b, t, 0].
//   prune_adjust_kernel  one CTA per utterance: padding fix-up + two suffix-min scans
//            (k2.monotonic_lower_bound) with the +-(R-1)*t transform between them in shared memory, then
//            the coalesced int64 store of ranges[b, t, 0..R).
// HBM-bound integer/compare work: reads (S*(T+1) + (S+1)*T)*4 B, writes T*R*8 B per utterance.
#include "common.cuh"

namespace s2t {
namespace {

constexpr int kThreads = 256;

// suffix-min over sm[0..T) in place (Hillis-Steele, chunked from the right)
__device__ void suffix_min_inplace(int* sm, int* tmp, int T) {
  // process chunks of blockDim from the right end carrying the running min
  __shared__ int carry;
  if (threadIdx.x == 0) carry = INT_MAX;
  __syncthreads();
  const int n = blockDim.x;
  for (int hi = T; hi > 0; hi -= n) {
    const int lo = max(hi - n, 0);
    const int len = hi - lo;
    const int i = threadIdx.x;
    int v = (i < len) ? sm[lo + i] : INT_MAX;
    // inclusive suffix scan inside the chunk
    for (int off = 1; off < n; off <<= 1) {
      tmp[i] = v;
      __syncthreads();
      if (i + off < len) v = min(v, tmp[i + off]);
      __syncthreads();
    }
    const int c = carry;
    __syncthreads();
    if (i < len) {
      v = min(v, c);
      sm[lo + i] = v;
    }
    __syncthreads();
    if (i == 0) carry = v;  // element lo holds the min of everything to its right
    __syncthreads();
  }
}

constexpr int kSegs = 8;

template <int VARIANT>
__global__ void __launch_bounds__(32 * kSegs)
prune_argmax_kernel(const float* __restrict__ px_grad, const float* __restrict__ py_grad,
                    const int64_t* __restrict__ boundary, int S, int T, int R, int64_t* __restrict__ ranges) {
  __shared__ float seg_v[kSegs][32];
  __shared__ int seg_s[kSegs][32];
  const int b = blockIdx.y;
  const int t = blockIdx.x * 32 + threadIdx.x;
  const int seg = threadIdx.y;
  const int S1 = S + 1, T1 = T + 1;
  const float* pxg = px_grad + (int64_t)b * S * T1;
  const float* pyg = py_grad + (int64_t)b * S1 * T;
  const int Sb = (int)boundary[4 * b + 2], Tb = (int)boundary[4 * b + 3];
  const int ncand = S1 - R + 1;
  const int per = (ncand + kSegs - 1) / kSegs;
  const int s0 = VARIANT == 0 ? min(seg * per, ncand) : 0, s1 = VARIANT == 0 ? min(s0 + per, ncand) : 0;
  float best_v = kNegInf;
  int best = s0;
  if (t < Tb - 1 && t < T && s0 < s1) {
    if (VARIANT == 0) {
      // The window of R occupation values slides through registers (one new py_grad load + one px_grad
      // load per candidate); the sum is still formed left to right over the window, which is the
      // bit-exactness contract.
      constexpr int kMaxR = 8;
      if (R <= kMaxR) {
        float w[kMaxR];
#pragma unroll
        for (int k = 0; k < kMaxR; ++k) w[k] = (k < R - 1) ? __ldg(pyg + (int64_t)(s0 + k) * T + t) : 0.f;
        const float* py_in = pyg + (int64_t)(R - 1) * T + t;
        const float* px_in = pxg + t - T1;  // px_grad[s-1, t]
#pragma unroll 4
        for (int s = s0; s < s1; ++s) {
          const float newest = __ldg(py_in + (int64_t)s * T);
          const float pxp = (s == 0) ? 0.f : __ldg(px_in + (int64_t)s * T1);
          float acc = w[0];
#pragma unroll
          for (int k = 1; k < kMaxR; ++k) {
            if (k < R - 1) acc += w[k];
          }
          if (R > 1) acc += newest; else acc = newest;
          const float v = acc - pxp;
          if (v > best_v) {
            best_v = v;
            best = s;
          }
#pragma unroll
          for (int k = 0; k < kMaxR - 1; ++k) w[k] = w[k + 1];
          if (R >= 2) {
#pragma unroll
            for (int k = 0; k < kMaxR; ++k)
              if (k == R - 2) w[k] = newest;
          }
        }
      } else {
        for (int s = s0; s < s1; ++s) {
          float acc = __ldg(pyg + (int64_t)s * T + t);
          for (int k = 1; k < R; ++k) acc += __ldg(pyg + (int64_t)(s + k) * T + t);
          float pxp = (s == 0) ? 0.f : __ldg(pxg + (int64_t)(s - 1) * T1 + t);
          const float v = acc - pxp;
          if (v > best_v) {
            best_v = v;
            best = s;
          }
        }
      }
    }
  }
  if (VARIANT == 1) {
    // B: cs[s+R] - cs[s] with cs the running sum over s of tot[s] = px_grad[s] + py_grad[s], formed sequentially
    // from 0 exactly as the oracle does (the contract leaves no freedom in the order of the sum).  What can run in
    // parallel are the loads: all kSegs threads of a frame fetch tot[] into shared memory (coalesced along t, eight
    // rows in flight per frame), then one thread per frame walks the two pointers cs_lo / cs_hi over it.
    extern __shared__ float tot[];  // [S1][32]
    if (t < T) {
      for (int j = seg; j < S1; j += kSegs)
        tot[j * 32 + threadIdx.x] = (j < S ? __ldg(pxg + (int64_t)j * T1 + t) : 0.f) + __ldg(pyg + (int64_t)j * T + t);
    }
    __syncthreads();
    if (seg == 0 && t < Tb - 1 && t < T) {
      const float* col = tot + threadIdx.x;
      float cs_hi = 0.f;
      for (int j = 0; j < R; ++j) cs_hi += col[j * 32];
      float cs_lo = 0.f;
#pragma unroll 4
      for (int s = 0; s < ncand; ++s) {
        const float v = cs_hi - cs_lo;
        if (v > best_v) {
          best_v = v;
          best = s;
        }
        if (s + 1 < ncand) {
          cs_lo += col[s * 32];
          cs_hi += col[(s + R) * 32];
        }
      }
    }
  }
  seg_v[seg][threadIdx.x] = best_v;
  seg_s[seg][threadIdx.x] = best;
  __syncthreads();
  if (seg == 0 && t < T) {
    int out;
    if (t < Tb - 1) {
      out = 0;
      float bv = kNegInf;
#pragma unroll
      for (int g = 0; g < kSegs; ++g) {
        if (seg_v[g][threadIdx.x] > bv) {
          bv = seg_v[g][threadIdx.x];
          out = seg_s[g][threadIdx.x];
        }
      }
    } else {
      out = max(Sb - R + 1, 0);  // last real frame and padding frames
    }
    ranges[((int64_t)b * T + t) * R] = out;
  }
}

__global__ void __launch_bounds__(kThreads)
prune_adjust_kernel(int T, int R, int64_t* __restrict__ ranges) {
  extern __shared__ int sm[];  // T ints + blockDim scratch
  int* sbeg = sm;
  int* tmp = sm + T;
  int64_t* out = ranges + (int64_t)blockIdx.x * T * R;
  for (int t = threadIdx.x; t < T; t += blockDim.x) sbeg[t] = (int)out[(int64_t)t * R];
  __syncthreads();
  // _adjust_pruning_lower_bound
  suffix_min_inplace(sbeg, tmp, T);
  for (int t = threadIdx.x; t < T; t += blockDim.x) sbeg[t] = -(sbeg[t] - (R - 1) * t);
  __syncthreads();
  suffix_min_inplace(sbeg, tmp, T);
  for (int t = threadIdx.x; t < T; t += blockDim.x) sbeg[t] = -(max(sbeg[t], 0) - (R - 1) * t);
  __syncthreads();
  for (int i = threadIdx.x; i < T * R; i += blockDim.x) out[i] = (int64_t)sbeg[i / R] + (i % R);
}

}  // namespace

int prune_ranges(const float* px_grad, const float* py_grad, const int64_t* boundary, int B, int S,
                 int T, int R, int variant, int64_t* ranges, cudaStream_t stream) {
  S2T_REQUIRE(R >= 2 && R <= S + 1, "prune_ranges: s_range=%d must be in [2, S+1=%d]", R, S + 1);
  S2T_REQUIRE(variant == 0 || variant == 1, "prune_ranges: unknown variant %d", variant);
  if (B == 0 || T == 0) return 0;
  size_t smem = (size_t)(T + kThreads) * sizeof(int);
  S2T_REQUIRE(smem <= 200 * 1024, "prune_ranges: T=%d too long for one CTA", T);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(prune_adjust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  ProfScope prof("prune_ranges_kernel", stream, 2);
  const dim3 grid((unsigned)((T + 31) / 32), (unsigned)B), block(32, kSegs);
  if (variant == 0) {
    prune_argmax_kernel<0><<<grid, block, 0, stream>>>(px_grad, py_grad, boundary, S, T, R, ranges);
  } else {
    const size_t smem_b = (size_t)(S + 1) * 32 * sizeof(float);
    S2T_REQUIRE(smem_b <= 200 * 1024, "prune_ranges: S=%d too long for the cumulative variant's shared-memory column tile", S);
    if (smem_b > 48 * 1024)
      cudaFuncSetAttribute(prune_argmax_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    prune_argmax_kernel<1><<<grid, block, smem_b, stream>>>(px_grad, py_grad, boundary, S, T, R, ranges);
  }
  prune_adjust_kernel<<<B, kThreads, smem, stream>>>(T, R, ranges);
  return check_launch("prune_ranges_kernel");
}

}  // namespace s2t
