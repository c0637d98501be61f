// Shared helpers for the sm_100a kernels of the transducer-loss path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

namespace s2t {

constexpr float kNegInf = -INFINITY;
// k2's LogAdd cut-off: log(FLT_EPSILON)  
constexpr float kMinLogDiff = -15.942385152878742f;

// Error plumbing for the C ABI (thread-local message, integer code).
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Counts kernel launches and, when bench.py enabled profiling, times them with CUDA events.
class ProfScope {
 public:
  ProfScope(const char* name, cudaStream_t stream, int launches = 1);
  ~ProfScope();
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;

 private:
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t start_, stop_;
  bool on_;
};

// Fork / join of independent kernel chains inside one library call: the caller's stream forks into up to two
// library-owned side streams (events only, no host synchronisation, legal under CUDA-graph capture) so that the tail
// of one persistent kernel overlaps the start of an independent one.  S2T_B200_NO_FORK=1 keeps everything on the
// caller's stream.
class ForkJoin {
 public:
  explicit ForkJoin(cudaStream_t main);
  cudaStream_t side(int i);  // i in {0, 1}; forks on first use
  void join();               // the main stream waits for every side stream that was used
  ~ForkJoin() { join(); }

 private:
  cudaStream_t main_;
  bool used_[2];
  bool enabled_;
};

// Per-device facts the launchers size their grids and scratch buffers from (queried once per device, cached).
struct DeviceInfo {
  int device;
  int sms;            // multiprocessors
  size_t total_mem;   // bytes of device memory
};
const DeviceInfo& device_info();  // of the CURRENT device

// Zero-fill of the hot path's accumulation buffers as a kernel (16-byte stores); S2T_B200_MEMSET_NODES=1 uses
// cudaMemsetAsync instead (measured equal inside the step's CUDA graph: 0.9326 vs 0.9334 ms).
void zero_async(void* p, size_t bytes, cudaStream_t stream);
constexpr int kMaxDevices = 64;

#define S2T_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::s2t::set_error(__VA_ARGS__);      \
      return 1;                           \
    }                                     \
  } while (0)

__device__ __forceinline__ float log_add(float x, float y) {
  // Same control flow as k2's LogAdd: NaN diff (-inf, -inf) returns the max.
  float diff;
  if (x < y) {
    diff = x - y;
    x = y;
  } else {
    diff = y - x;
  }
  if (diff >= kMinLogDiff) return x + log1pf(expf(diff));
  return x;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half(v); }

enum Activation : int { kRelu = 0, kTanh = 1 };

__device__ __forceinline__ float act_fwd(float x, int act) {
  return act == kRelu ? fmaxf(x, 0.f) : tanhf(x);
}
// derivative expressed through the pre-activation value x
__device__ __forceinline__ float act_bwd(float x, int act) {
  if (act == kRelu) return x > 0.f ? 1.f : 0.f;
  float y = tanhf(x);
  return 1.f - y * y;
}

}  // namespace s2t
