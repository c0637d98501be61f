// Launch accounting and an opt-in per-kernel CUDA-event timer for bench.py.
//
// Every kernel launch of the library goes through a ProfScope: it bumps the
// launch counter (bench.py reports it as "gpu_launches") and, while profiling is
// enabled, brackets the launch with two events on the launching stream so that
// bench.py can attribute device time to kernels inside its own timed region
// (no external profiler attached).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <stdio.h>

#include "common.cuh"
#include <stdlib.h>

namespace s2t {

namespace {
struct Entry {
  const char* name;
  cudaEvent_t start, stop;
};
std::atomic<long long> g_launches{0};
std::atomic<bool> g_enabled{false};
std::mutex g_mu;
std::vector<Entry> g_entries;
}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t stream, int launches) : name_(name), stream_(stream), on_(false) {
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (g_enabled.load(std::memory_order_relaxed)) {
    on_ = true;
    cudaEventCreate(&start_);
    cudaEventCreate(&stop_);
    cudaEventRecord(start_, stream_);
  }
}

ProfScope::~ProfScope() {
  if (!on_) return;
  cudaEventRecord(stop_, stream_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_entries.push_back(Entry{name_, start_, stop_});
}

// ---- DeviceInfo ----------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) zero_kernel(uint4* __restrict__ p, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
}  // namespace

void zero_async(void* p, size_t bytes, cudaStream_t stream) {
  static const bool nodes = getenv("S2T_B200_MEMSET_NODES") != nullptr;
  if (bytes == 0) return;
  if (nodes || (reinterpret_cast<uintptr_t>(p) & 15) != 0 || (bytes & 15) != 0) {
    cudaMemsetAsync(p, 0, bytes, stream);
    return;
  }
  const size_t n16 = bytes / 16;
  size_t blocks = (n16 + 255) / 256;
  const size_t cap = (size_t)device_info().sms * 8;
  if (blocks > cap) blocks = cap;
  zero_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(p), n16);
}

const DeviceInfo& device_info() {
  static DeviceInfo info[kMaxDevices];
  static std::atomic<int> ready[kMaxDevices];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= (kMaxDevices - 1);
  if (ready[dev].load(std::memory_order_acquire) == 0) {
    std::lock_guard<std::mutex> lk(mu);
    if (ready[dev].load(std::memory_order_relaxed) == 0) {
      DeviceInfo d{};
      d.device = dev;
      cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
      if (d.sms <= 0) d.sms = 148;
      size_t free_b = 0, total_b = 0;
      d.total_mem = (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) ? total_b : ((size_t)32 << 30);
      info[dev] = d;
      ready[dev].store(1, std::memory_order_release);
    }
  }
  return info[dev];
}

// ---- ForkJoin -------------------------------------------------------------------------------------
namespace {
struct SideStreams {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, done[2] = {nullptr, nullptr};
  bool ok = false;
};
// One pair of side streams per (device, caller stream): two independent chains of the step that run on different
// caller streams (the simple-loss gradients next to the joiner forward, functional._SimpleLoss) must not meet on a
// shared side stream, where the second chain would queue behind the first one's kernels.
SideStreams& side_streams(cudaStream_t main) {
  static thread_local std::map<std::pair<int, cudaStream_t>, SideStreams> table;
  int dev = 0;
  cudaGetDevice(&dev);
  SideStreams& st = table[std::make_pair(dev, main)];
  if (!st.ok) {
    for (int i = 0; i < 2; ++i) {
      cudaStreamCreateWithFlags(&st.s[i], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&st.done[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&st.fork, cudaEventDisableTiming);
    st.ok = true;
  }
  return st;
}
}  // namespace

ForkJoin::ForkJoin(cudaStream_t main) : main_(main), used_{false, false} {
  static const bool disabled = getenv("S2T_B200_NO_FORK") != nullptr;
  enabled_ = !disabled && !g_enabled.load();  // the per-kernel timer wants exclusive kernel times: no overlap
}

cudaStream_t ForkJoin::side(int i) {
  if (!enabled_) return main_;
  SideStreams& st = side_streams(main_);
  if (!used_[i]) {
    cudaEventRecord(st.fork, main_);
    cudaStreamWaitEvent(st.s[i], st.fork, 0);
    used_[i] = true;
  }
  return st.s[i];
}

void ForkJoin::join() {
  if (!enabled_) return;
  SideStreams& st = side_streams(main_);
  for (int i = 0; i < 2; ++i) {
    if (!used_[i]) continue;
    cudaEventRecord(st.done[i], st.s[i]);
    cudaStreamWaitEvent(main_, st.done[i], 0);
    used_[i] = false;
  }
}

}  // namespace s2t

using namespace s2t;

extern "C" {

long long s2t_launch_count(void) { return g_launches.load(); }

void s2t_profile_enable(int on) { g_enabled.store(on != 0); }

// Waits for the recorded events, writes "name\tlaunch_groups\ttotal_ms\n" lines into buf
// (truncated to n bytes) and clears the record.  Returns the number of distinct names.
int s2t_profile_report(char* buf, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& e : g_entries) {
    cudaEventSynchronize(e.stop);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.start, e.stop);
    auto& a = agg[e.name];
    a.first += 1;
    a.second += ms;
    cudaEventDestroy(e.start);
    cudaEventDestroy(e.stop);
  }
  g_entries.clear();
  size_t off = 0;
  if (n > 0) buf[0] = 0;
  for (auto& kv : agg) {
    int w = snprintf(buf + off, off < n ? n - off : 0, "%s\t%d\t%.6f\n", kv.first.c_str(), kv.second.first,
                     kv.second.second);
    if (w < 0 || off + (size_t)w >= n) break;
    off += (size_t)w;
  }
  return (int)agg.size();
}

}  // extern "C"
